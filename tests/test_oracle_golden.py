"""Pin the oracle: the reference hot path compiled in place (oracle/_ref) must reproduce the
reference's own golden vectors, and the committed fixtures must agree with it."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

REF_TESTS = "/root/reference/src/fvm/test/"


def test_testLinearSolver_stdout_is_byte_identical(ref):
    exe = os.path.join(ROOT, "oracle", "_ref", "testLinearSolver")
    if not (os.path.exists(exe) and os.path.isdir(REF_TESTS)):
        pytest.skip("reference tree not mounted (GPU box)")
    out = subprocess.run([exe, "MatrixMarket226.dat", "rhs226.dat"], cwd=REF_TESTS, capture_output=True, text=True)
    assert out.stdout == open(REF_TESTS + "testLinearSolver.out").read()


def test_mm226_fixture_matches_reference_golden(ref):
    g = load_golden("mm226.npz")
    # T/testLinearSolver.out:5-11 -- levels 108/48/20/8/3, 6981.57 -> 5.32223e-05 at cycle 40
    assert g["ref_levels"].tolist() == [108, 48, 20, 8, 3]
    text = str(g["golden_text"])
    assert "0: [test : 6981.57]" in text and "40: [test : 5.32223e-05]" in text
    r = ref.linsolve(int(g["n"]), g["row"], g["col"], g["diag"], g["off"], g["b"], ref.solver_cfg(verbosity=2))
    assert r["levels"] == [108, 48, 20, 8, 3] and r["iters"] == 40
    assert "40: [test : 5.32223e-05]" in r["text"]
    assert np.array_equal(r["x"], g["ref_x_tol8"])


def test_cav32_thermal_matrix_golden(ref):
    """T/THERMAL_MATRIX/GOLDEN/{matrix_mesh0.mat,matrix.rhs} and the AMG history of
    T/AMG_MERGING_THERMAL/proc1/GOLDEN/convergence.dat, from the fixture's mesh arrays alone."""
    g = load_golden("cav32.npz")
    n = int(g["n_self"])
    assert np.abs(-g["b"][:n] - g["golden_rhs"]).max() == 0.0
    row, col = g["cc_row"], g["cc_col"]
    ents = []
    for i in range(n):
        ents.append((i + 1, i + 1, g["diag"][i]))
        for jp in range(row[i], row[i + 1]):
            if col[jp] < n:
                ents.append((i + 1, col[jp] + 1, g["off"][jp]))
    assert np.abs(np.array(ents) - g["golden_mat"]).max() == 0.0
    conv = str(g["golden_convergence"]).splitlines()
    assert conv[0] == "0: [therm.temperature : 63200]" and conv[1] == "56: [therm.temperature : 5.75812e-05]"
    assert conv[0] in str(g["ref_text"]) and conv[1] in str(g["ref_text"])


def test_cav32_live_reference_matches_fixture(ref):
    if not os.path.isdir(REF_TESTS):
        pytest.skip("reference tree not mounted (GPU box)")
    g = load_golden("cav32.npz")
    rm = ref.RefMesh.from_cas(REF_TESTS + "cav32.cas")
    t = ref.RefThermal(rm)
    t.set_bc(3, "SpecifiedTemperature", specifiedTemperature=400)
    for gid in (4, 5, 6):
        t.set_bc(gid, "SpecifiedTemperature", specifiedTemperature=0)
    t.set_solver(ref.solver_cfg(relativeTolerance=1e-9, nMaxIterations=2000, maxCoarseLevels=20, verbosity=2))
    t.init()
    a = t.assemble(1)
    for k, gk in (("diag", "diag"), ("offdiag", "off"), ("b", "b")):
        assert np.array_equal(a[k], g[gk])
    text, _ = t.advance(1)
    assert "56: [therm.temperature : 5.75812e-05]" in text
    assert np.array_equal(t.field("temperature"), g["ref_x_tol9"])


def test_mesh_generators_match_reference_metrics(ref):
    from fvm_b200 import meshgen as G
    for m in (G.quad_mesh(7, 5, jitter=0.2), G.hex_mesh(4, 5, 3, jitter=0.2), G.tet_mesh(3, 4, 2)):
        rm = ref.RefMesh.from_raw(m.dim, m.n_cells, m.nodes, m.face_cells, m.face_nodes, m.face_node_count,
                                  m.face_group_size)
        c, geo = rm.connectivity(), rm.geometry()
        row, col = G.connectivity(m)
        assert np.array_equal(row, c["cc_row"]) and np.array_equal(col, c["cc_col"])
        mt = G.metrics(m)
        for k in mt:
            assert np.abs(mt[k] - geo[k]).max() < 1e-13, k
        assert geo["cell_volume"].min() > 0
