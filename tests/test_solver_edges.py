"""Edge cases of the linear-solver path (degenerate sizes, empty right-hand sides, iteration
limits, disconnected patterns), the solver family beside AMG (ILU0, BCGStab + ILU0, JacobiSolver) and the
reference-order verification mode. Every test runs twice: on the B200 through the C ABI (-m gpu) and in the
kernels' host simulator (index logic on the GPU-less box)."""
import numpy as np
import pytest

from fvm_b200 import capi as X


def raw_system(lib, n, edges, diag, b, vals=None):
    adj = [[] for _ in range(n)]
    for k, (i, j) in enumerate(edges):
        v = -1.0 if vals is None else vals[k]
        adj[i].append((j, v)); adj[j].append((i, v))
    row = np.zeros(n + 1, np.int32)
    row[1:] = np.cumsum([len(a) for a in adj])
    col = np.array([j for a in adj for j, _ in a] or [0], np.int32)[: row[-1]] if row[-1] else np.zeros(0, np.int32)
    off = np.array([v for a in adj for _, v in a], float)
    return X.DeviceSystem(lib, raw=(n, 0, row, col, np.asarray(diag, float), off, np.asarray(b, float)))


def residual(n, edges, diag, b, x, vals=None):
    r = np.asarray(b, float) + np.asarray(diag, float) * x
    for k, (i, j) in enumerate(edges):
        v = -1.0 if vals is None else vals[k]
        r[i] += v * x[j]; r[j] += v * x[i]
    return r


def test_single_row_system(devlib):
    ds = raw_system(devlib, 1, [], [4.0], [2.0])
    amg = X.DeviceAMG(devlib)
    r0, r, it = amg.solve(ds)
    assert r0 == 2.0 and it >= 1 and r <= 1e-8 * r0
    assert ds.get_field(X.FIELD_DELTA)[0] == -0.5        # A delta + b = 0
    amg.close(); ds.close()


def test_zero_right_hand_side_needs_no_cycle(devlib):
    """AMG::solve returns at once when the initial residual is below the absolute tolerance (F/AMG.cpp:240)."""
    edges = [(i, i + 1) for i in range(9)]
    ds = raw_system(devlib, 10, edges, np.full(10, 2.5), np.zeros(10))
    amg = X.DeviceAMG(devlib)
    r0, r, it = amg.solve(ds)
    assert r0 == 0.0 and it == 0 and not ds.get_field(X.FIELD_DELTA).any()
    amg.close(); ds.close()


def test_iteration_limit_is_respected(devlib):
    """for (i = 1; i < nMaxIterations; i++) -- F/AMG.cpp:245: at most nMaxIterations - 1 cycles."""
    n = 400
    edges = [(i, i + 1) for i in range(n - 1)]
    diag = np.full(n, 2.0); diag[0] = 3.0
    ds = raw_system(devlib, n, edges, diag, np.ones(n))
    o = devlib.default_amg_opts()
    o.nMaxIterations, o.relativeTolerance = 4, 1e-30
    amg = X.DeviceAMG(devlib, o)
    r0, r, it = amg.solve(ds)
    assert it == 3 and 0 < r < r0 and len(amg.history()) == 4
    amg.close(); ds.close()


def test_diagonal_matrix_is_one_colour_and_one_cycle(devlib):
    n = 50
    rng = np.random.default_rng(1)
    diag, b = rng.uniform(1, 3, n), rng.normal(size=n)
    ds = raw_system(devlib, n, [], diag, b)
    amg = X.DeviceAMG(devlib)
    r0, r, it = amg.solve(ds)
    assert it == 1 and amg.levels()["colours"][0] == 1
    assert np.allclose(ds.get_field(X.FIELD_DELTA), -b / diag, rtol=1e-15, atol=0)
    amg.close(); ds.close()


def test_disconnected_blocks_and_isolated_rows(devlib):
    """Two chains, a ring of odd length (not bipartite) and three isolated rows in one system."""
    edges = [(i, i + 1) for i in range(0, 19)] + [(i, i + 1) for i in range(20, 39)]
    ring = list(range(40, 47))
    edges += [(ring[k], ring[(k + 1) % 7]) for k in range(7)]
    n = 50
    rng = np.random.default_rng(2)
    diag = np.full(n, 2.2)
    b = rng.normal(size=n)
    ds = raw_system(devlib, n, edges, diag, b)
    o = devlib.default_amg_opts()
    o.relativeTolerance, o.nMaxIterations = 1e-13, 500
    amg = X.DeviceAMG(devlib, o)
    r0, r, it = amg.solve(ds)
    x = ds.get_field(X.FIELD_DELTA)
    assert np.abs(residual(n, edges, diag, b, x)).sum() <= 1e-12 * r0
    assert amg.levels()["colours"][0] == 3           # the odd ring needs a third class
    amg.close(); ds.close()


@pytest.mark.parametrize("levels", [0, 1, 3])
def test_max_coarse_levels_is_a_hard_cap(devlib, levels):
    """AMG::maxCoarseLevels (F/AMG.cpp:154): 0 = smoothing on the fine level only."""
    n = 256
    edges = [(i, i + 1) for i in range(n - 1)]
    diag = np.full(n, 2.0); diag[0] = 3.0; diag[-1] = 3.0
    ds = raw_system(devlib, n, edges, diag, np.ones(n))
    o = devlib.default_amg_opts()
    o.maxCoarseLevels, o.nMaxIterations, o.relativeTolerance = levels, 30, 1e-30
    amg = X.DeviceAMG(devlib, o)
    r0, r, it = amg.solve(ds)
    assert len(amg.levels()["sizes"]) == levels + 1 and r < r0
    amg.close(); ds.close()


def test_w_and_f_cycles_converge_in_fewer_cycles_than_v(devlib):
    n = 900
    edges = [(y * 30 + x, y * 30 + x + 1) for y in range(30) for x in range(29)]
    edges += [(y * 30 + x, (y + 1) * 30 + x) for y in range(29) for x in range(30)]
    diag = np.full(n, 4.0 + 1e-3)
    its = {}
    for name, ct in (("V", X.CYCLE_V), ("F", X.CYCLE_F), ("W", X.CYCLE_W)):
        ds = raw_system(devlib, n, edges, diag, np.ones(n))
        o = devlib.default_amg_opts()
        o.cycleType, o.relativeTolerance, o.nMaxIterations = ct, 1e-10, 2000
        amg = X.DeviceAMG(devlib, o)
        r0, r, it = amg.solve(ds)
        assert r / r0 < 1e-10
        its[name] = it
        amg.close(); ds.close()
    assert its["W"] <= its["F"] <= its["V"]


# ---------------------------------------------------------------- ILU(0): ILU0Solver and BCGStab + ILU0
def _cav32_raw(lib):
    from conftest import load_golden
    g = load_golden("cav32.npz")
    n, nt = int(g["n_self"]), int(g["n_total"])
    return g, n, nt, X.DeviceSystem(lib, raw=(n, nt - n, g["cc_row"], g["cc_col"], g["diag"], g["off"], g["b"]))


def test_ilu0_factors_and_solves_are_bit_identical_to_the_reference(devlib):
    """CRMatrix::compute_ILU0 / lowerSolve / upperSolve (F/CRMatrix.h:1546-1715) row by row in dependency levels:
    ILU0Solver's delta on the cav32 conduction matrix equals the reference's bit for bit, and so does the
    residual after a sweep (a second sweep recomputes the same delta, as in the reference)."""
    from oracle import refapi as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    g, n, nt, ds = _cav32_raw(devlib)
    amg = X.DeviceAMG(devlib)
    r0, r, it = amg.ilu0(ds, 5, 1e-9)
    ref = R.linsolve(n, g["cc_row"], g["cc_col"], g["diag"], g["off"], g["b"],
                     R.solver_cfg(kind=2, nMaxIterations=5, verbosity=1, relativeTolerance=1e-9), n_ghost=nt - n)
    assert r0 == ref["rnorm0"] == 63200.0 and it == 4
    assert np.array_equal(ds.get_field(X.FIELD_DELTA)[:n], ref["x"][:n])
    assert "%g" % r == ref["text"].splitlines()[-1].split(":")[-1].strip(" ]")
    assert amg.ilu_levels == 63          # 32 x 32 quads in natural order: 32 + 32 - 1 wavefronts
    h = amg.history()
    assert h[1] == h[2] == h[3]
    amg.close(); ds.close()


def test_bcgstab_with_ilu0_preconditioner_follows_the_reference_iteration_for_iteration(devlib):
    """T/PARALLEL_CAVITY_ILU0's solver pairing on the cav32 matrix: same 28 iterations, same residual history
    (to 1e-4: the reference prints 6 digits and the dot products are reduced in a different order), same
    solution (1e-12)."""
    from oracle import refapi as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    g, n, nt, ds = _cav32_raw(devlib)
    amg = X.DeviceAMG(devlib)
    r0, r, it = amg.bcgstab_ilu0(ds, 200, 1e-9)
    ref = R.linsolve(n, g["cc_row"], g["cc_col"], g["diag"], g["off"], g["b"],
                     R.solver_cfg(kind=3, nMaxIterations=200, verbosity=1, relativeTolerance=1e-9), n_ghost=nt - n)
    lines = [l for l in ref["text"].splitlines() if l[0].isdigit()]
    assert it == int(lines[-1].split(":")[0]) == 28
    hist = amg.history()
    ref_hist = np.array([float(l.split(":")[-1].strip(" ]")) for l in lines])
    assert len(hist) == len(ref_hist) and np.abs(hist / ref_hist - 1.0).max() < 1e-4   # 6 printed digits, dots reduced in another order
    x = ds.get_field(X.FIELD_DELTA)
    assert np.abs(x[:n] - ref["x"][:n]).max() <= 1e-12 * np.abs(ref["x"][:n]).max()
    # an AMG solve on the same handle afterwards still builds its full hierarchy
    ds2 = _cav32_raw(devlib)[3]
    amg.solve(ds2)
    assert len(amg.levels()["sizes"]) > 5
    amg.close(); ds.close(); ds2.close()


def test_ilu0_solver_classes_in_the_model_api(devlib):
    """fvmbaseExt.ILU0Solver as BCGStab's preconditioner in ThermalModelA (the reference's script pattern)."""
    from fvm_b200 import meshgen as G, models as M
    import contextlib
    import io
    raw = G.quad_mesh(20, 16, jitter=0.15, seed=2)
    out = {}
    for name in ("amg", "ilu0"):
        mesh = M.Mesh(raw)
        geom = M.GeomFields("geom")
        M.MeshMetricsCalculatorA(geom, [mesh], lib=devlib).init()
        tf = M.ThermalFields("therm")
        tm = M.ThermalModelA(geom, tf, [mesh], lib=devlib)
        bc = tm.getBCMap()
        bc[3].bcType = "SpecifiedTemperature"; bc[3].setVar("specifiedTemperature", 400)
        bc[4].bcType = "SpecifiedTemperature"; bc[4].setVar("specifiedTemperature", 300)
        ls = M.BCGStab()
        ls.preconditioner = M.AMG() if name == "amg" else M.ILU0Solver()
        ls.relativeTolerance, ls.nMaxIterations, ls.verbosity = 1e-13, 500, 0
        tm.getOptions().linearSolver = ls
        tm.init()
        with contextlib.redirect_stdout(io.StringIO()):
            tm.advance(1)
        out[name] = tf.temperature[mesh.getCells()].copy()
    assert np.abs(out["amg"] - out["ilu0"]).max() <= 1e-9 * np.abs(out["amg"]).max()


def test_jacobi_solver_is_bit_identical_to_the_reference_class(devlib):
    """F/JacobiSolver.cpp on T/MatrixMarket226.dat: 1265 iterations to rel 1e-12 in the reference and here, the
    same last residual to the last digit and the same delta bit for bit (every row's sum runs in entry order)."""
    from conftest import load_golden
    from oracle import refapi as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    g = load_golden("mm226.npz")
    n = int(g["n"])
    ds = X.DeviceSystem(devlib, raw=(n, 0, g["row"], g["col"], g["diag"], g["off"], g["b"]))
    amg = X.DeviceAMG(devlib)
    r0, r, it = amg.jacobi(ds, 20000, 1e-12, 1e-50)
    ref = R.linsolve(n, g["row"], g["col"], g["diag"], g["off"], g["b"],
                     R.solver_cfg(kind=5, nMaxIterations=20000, verbosity=1, relativeTolerance=1e-12))
    last = ref["text"].splitlines()[-1]
    assert it == int(last.split(":")[0]) == 1265
    ref_r = float(last.split(":")[-1].strip(" ]"))
    if "hostsim" in devlib.path:
        assert "%g" % r == "%g" % ref_r
        assert np.array_equal(ds.get_field(X.FIELD_DELTA), ref["x"])
    else:   # the device's row sums contract a*x+s into FMAs: same iterates to rounding, not to the bit -- and the
        # last residual (1e-12 of the first) is itself made of rounding errors: same size, not the same digits
        assert abs(r - ref_r) <= 1e-3 * ref_r
        assert np.abs(ds.get_field(X.FIELD_DELTA) - ref["x"]).max() <= 1e-12 * np.abs(ref["x"]).max()
    amg.close(); ds.close()


# ---------------------------------------------------------------- reference-order verification mode
def test_reference_order_reproduces_testLinearSolver_golden(devlib, reference_order_dev):
    """T/TESTS Fvm001 with this library's own solver: in reference-order mode (sequential greedy agglomeration on the
    host, Gauss-Seidel scheduled by the dependency levels of the natural numbering) the hierarchy is the reference's
    -- 226 -> 108 / 48 / 20 / 8 / 3 rows -- and the V-cycles are the reference's: 40 cycles, last residual 5.32223e-05
    as in testLinearSolver.out, solution equal to the reference's to rounding (1e-15)."""
    from conftest import load_golden
    g = load_golden("mm226.npz")
    n = int(g["n"])
    ds = X.DeviceSystem(devlib, raw=(n, 0, g["row"], g["col"], g["diag"], g["off"], g["b"]))
    amg = X.DeviceAMG(devlib)
    r0, r, it = amg.solve(ds)
    assert amg.levels()["sizes"] == [226] + [int(v) for v in g["ref_levels"]]
    assert it == int(g["ref_iters"]) == 40
    golden = [l for l in str(g["golden_text"]).splitlines() if l and l[0].isdigit()]
    assert golden == ["0: [test : %g]" % r0, "%d: [test : %g]" % (it, r)]
    x = ds.get_field(X.FIELD_DELTA)
    assert np.abs(x - g["ref_x_tol8"]).max() <= 1e-14 * np.abs(g["ref_x_tol8"]).max()
    amg.close(); ds.close()


def test_reference_order_reproduces_the_cav32_amg_golden(devlib, reference_order_dev):
    """T/AMG_MERGING_THERMAL/proc1/GOLDEN/convergence.dat (cav32 conduction, AMG to rel 1e-9): 56 cycles, 5.75812e-05."""
    g, n, nt, ds = _cav32_raw(devlib)
    o = devlib.default_amg_opts()
    o.relativeTolerance, o.nMaxIterations, o.maxCoarseLevels = 1e-9, 2000, 20
    amg = X.DeviceAMG(devlib, o)
    r0, r, it = amg.solve(ds)
    conv = str(g["golden_convergence"]).splitlines()
    assert conv[:2] == ["0: [therm.temperature : %g]" % r0, "%d: [therm.temperature : %g]" % (it, r)]
    x = ds.get_field(X.FIELD_DELTA)
    ref_delta = g["ref_x_tol9"][:n] - g["x_after_bc"][:n]
    assert np.abs(x[:n] - ref_delta).max() <= 1e-10 * np.abs(ref_delta).max()
    amg.close(); ds.close()
