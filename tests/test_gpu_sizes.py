"""Parity at sizes where the production code path differs from the one the small fixtures reach: several levels above
the cooperative kernel's row limit (per-level launches with 16-bit column offsets + CUDA graph + fused coarse stretch in
ONE cycle), hierarchies of 15+ levels, several CTAs per colour class. The checker is the reference's own C++ (oracle/_ref, prebuilt, travels to
the GPU box) run on the same mesh: assembly 1e-12, converged fields 1e-8 (north_star).
The same code runs in the host simulator at toy sizes (-m "not gpu") so that the test bodies themselves stay green."""
import contextlib
import io

import numpy as np
import pytest

import bench
import bench_workloads as W


@pytest.fixture
def lib_and_gpu(devlib, request):
    return devlib, request.node.callspec.params["devlib"] == "gpu"


def test_thermal_hex_above_the_cooperative_row_limit(lib_and_gpu, ref, monkeypatch):
    """112^3 jittered hexes (1.4 M rows: levels 0-3 run as per-level launches, the levels from 88 K rows down in
    k_coop_vcycle / k_tail_vcycle), random conductivity, flux + Dirichlet boundaries: bench.py's own parity block."""
    lib, gpu = lib_and_gpu
    n = 112 if gpu else 10
    monkeypatch.setattr(bench, "MESH", "tet")      # skip the exact-solution check of the bench workload (no workload here)
    monkeypatch.setattr(bench, "KRYLOV", False)
    out = bench.parity_block(lib, 0, 1, np.zeros(1), np.zeros(1), 1.0, 1, n)
    oc = out["oracle_check"]
    assert oc["oracle"].startswith("reference")
    assert oc["assembly_max_rel_diff"]["diag"] <= 1e-12 and oc["assembly_max_rel_diff"]["b"] <= 1e-12, oc
    assert oc["solution_rel_l2"] <= 1e-8, oc
    assert oc["cycles"] <= 2 * oc["reference_cycles"], oc


def test_cavity_128_three_simple_iterations(lib_and_gpu, ref):
    """128^2 lid-driven cavity, 3 SIMPLE iterations with converged inner solves on both sides: momentum by AMG
    (the multi-RHS cycles), pressure correction (reference cell pinned) by BCGStab + AMG -- plain V(0,1) cycles
    contract this system by only 0.995 per cycle at 128^2, on either side, and do not reach 1e-13."""
    from fvm_b200 import models as M
    lib, gpu = lib_and_gpu
    n, mu = (128 if gpu else 16), 0.01
    _, r = W._cavity_reference(n, mu, 3, tight=True, pressure_kind=1)
    _, mesh, ff, fm = W._cavity_model(lib, n, mu)
    o = fm.getOptions()
    o.momentumLinearSolver = M.AMG()
    o.pressureLinearSolver = M.BCGStab()
    o.pressureLinearSolver.preconditioner = M.AMG()
    o.pressureLinearSolver.preconditioner.verbosity = 0
    for s in (o.momentumLinearSolver, o.pressureLinearSolver):
        s.relativeTolerance, s.nMaxIterations, s.verbosity = 1e-13, 3000, 0
    with contextlib.redirect_stdout(io.StringIO()):
        fm.advance(3)
    cells, nn = mesh.getCells(), r["n_cells"]
    v = np.asarray(ff.velocity[cells]).reshape(-1, 3)[:nn]
    vr = r["velocity"].reshape(-1, 3)[:nn]
    p = np.asarray(ff.pressure[cells])[:nn]
    assert np.linalg.norm(v - vr) / np.linalg.norm(vr) <= 1e-8
    assert np.linalg.norm(p - r["pressure"][:nn]) / np.linalg.norm(r["pressure"][:nn]) <= 1e-8


def test_electric_model_on_32cubed_tets(lib_and_gpu, ref):
    """32^3 x 6 jittered tetrahedra (196 608 cells, unstructured coarse levels with 6-10 colour classes): two time
    steps of ElectricModel (electrostatics + drift / transient charge transport) against the reference's model."""
    from fvm_b200 import meshgen as G
    lib, gpu = lib_and_gpu
    n = 32 if gpu else 4
    raw = G.tet_mesh(n, n, n, lx=W.E_BOX, ly=W.E_BOX, lz=W.E_BOX)
    _, r = W._electric_reference(raw, 2, 1e-13, kind=0)
    mesh, ef, em = W._electric_model(lib, raw, tol=1e-13, iters=500)
    for _ in range(2):
        with contextlib.redirect_stdout(io.StringIO()):
            em.advance(1)
        em.updateTime()
    cells, nc = mesh.getCells(), raw.n_cells
    pot = np.asarray(ef.potential[cells])[:nc]
    chg = np.asarray(ef.charge[cells])[:nc, 2]
    assert np.linalg.norm(pot - r["potential"][:nc]) / np.linalg.norm(r["potential"][:nc]) <= 1e-8
    assert np.linalg.norm(chg - r["charge"][:nc, 2]) / np.linalg.norm(r["charge"][:nc, 2]) <= 1e-8
