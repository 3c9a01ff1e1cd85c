"""Pin the C restatement (oracle/fvm_oracle.c): it must reproduce the reference's golden vectors and,
where the compiled reference is available, agree with it bit for bit (it restates the same
sequential algorithm, so even cycle counts and level sizes must match)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import port


def test_port_amg_reproduces_testLinearSolver_golden():
    g = load_golden("mm226.npz")
    r = port.solve(int(g["n"]), g["row"], g["col"], g["diag"], g["off"], g["b"])
    assert r["levels"] == [108, 48, 20, 8, 3]                      # T/testLinearSolver.out:5-9
    assert r["iters"] == 40
    assert "%g" % r["history"][0] == "6981.57" and "%g" % r["history"][-1] == "5.32223e-05"   # :10-11
    assert np.array_equal(r["x"], g["ref_x_tol8"])                  # bit-identical to the compiled reference
    t = port.solve(int(g["n"]), g["row"], g["col"], g["diag"], g["off"], g["b"],
                   port.amg_opts(nMaxIterations=500, relativeTolerance=1e-13))
    assert np.array_equal(t["x"], g["ref_x"])
    b = port.solve(int(g["n"]), g["row"], g["col"], g["diag"], g["off"], g["b"],
                   port.amg_opts(nMaxIterations=100, relativeTolerance=1e-13), bcgstab=True)
    assert np.abs(b["x"] - g["ref_x_bcgstab"]).max() <= 1e-12 * np.abs(g["ref_x"]).max()


def test_port_thermal_cav32_golden():
    g = load_golden("cav32.npz")
    pm = port.PortMesh(g)
    assert np.array_equal(pm.pair_to_col, g["pair_to_col"])
    bcs = {3: ("dirichlet", [400.0]), 4: ("dirichlet", [0.0]), 5: ("dirichlet", [0.0]), 6: ("dirichlet", [0.0])}
    a = pm.assemble(np.full(pm.n_total, 300.0), bcs)
    n = pm.n_self
    assert np.abs(-a["b"][:n] - g["golden_rhs"]).max() == 0.0         # T/THERMAL_MATRIX/GOLDEN/matrix.rhs
    assert np.array_equal(a["diag"], g["diag"]) and np.array_equal(a["offdiag"], g["off"]) and np.array_equal(a["b"], g["b"])
    # T/AMG_MERGING_THERMAL/proc1/GOLDEN/convergence.dat: 63200 -> 5.75812e-05 in 56 cycles
    r = port.solve(n, g["cc_row"], g["cc_col"], a["diag"], a["offdiag"], a["b"],
                   port.amg_opts(nMaxIterations=2000, relativeTolerance=1e-9, maxCoarseLevels=20),
                   n_ghost=pm.n_total - n, is_boundary=a["is_boundary"])
    assert r["iters"] == 56 and "%g" % r["history"][0] == "63200" and "%g" % r["history"][-1] == "5.75812e-05"
    x, _ = pm.post_solve(a, r["x"])
    assert np.array_equal(x, g["ref_x_tol9"])


@pytest.mark.parametrize("stage", [0, 1])
def test_port_every_bc_kind_matches_reference_fixture(stage):
    g = load_golden("hex_bcs.npz")
    pm = port.PortMesh(g)
    bcs = {1: ("dirichlet", [400.0]), 2: ("neumann", [25.0]), 3: ("convective", [3.0, 280.0]),
           4: ("radiative", [0.8, 250.0]), 5: ("mixed", [2.0, 0.5, 310.0]), 6: ("neumann", [0.0])}
    a = pm.assemble(g["x0"], bcs, diffusivity=g["k"], source=g["src"], eliminate_boundary=stage)
    p = "s%d_" % stage
    for k in ("diag", "offdiag", "b", "x", "is_boundary", "gradient"):
        assert np.array_equal(a[k], g[p + k]), k


def test_port_tet_solution_and_fluxes():
    g = load_golden("tet_solve.npz")
    pm = port.PortMesh(g)
    bcs = {5: ("dirichlet", [300.0]), 6: ("dirichlet", [400.0]), 1: ("neumann", [7.0]), 2: ("neumann", [0.0]),
           3: ("neumann", [0.0]), 4: ("neumann", [0.0])}
    a0 = pm.assemble(np.full(pm.n_total, 300.0), bcs, diffusivity=g["k"])
    assert np.array_equal(a0["diag"], g["diag"]) and np.array_equal(a0["b"], g["b"])
    a = pm.assemble(a0["x"], bcs, diffusivity=g["k"])   # the fixture's advance() re-assembles
    r = port.solve(pm.n_self, g["cc_row"], g["cc_col"], a["diag"], a["offdiag"], a["b"],
                   port.amg_opts(nMaxIterations=2000, relativeTolerance=1e-13), n_ghost=pm.n_total - pm.n_self,
                   is_boundary=a["is_boundary"])
    x, bflux = pm.post_solve(a, r["x"])
    assert np.array_equal(x, g["ref_x"])
    for gi in range(1, 7):
        off, cnt = int(g["group_offset"][gi]), int(g["group_count"][gi])
        assert np.array_equal(bflux[off:off + cnt], g["hf%d" % int(g["group_id"][gi])])
