"""The host-side mirror of the reference's Python API: a script that reads like
T/THERMAL_MATRIX/testThermalParallel.py (main block) must run unchanged against the GPU path."""
import io
import contextlib
import os

import numpy as np
import pytest

from conftest import load_golden
from fvm_b200 import meshgen as G, models as M
from helpers import rel_l2


def build_model(lib, raw, tol=1e-9, solver="amg"):
    meshes = [M.Mesh(raw)]
    geomFields = M.GeomFields("geom")
    metricsCalculator = M.MeshMetricsCalculatorA(geomFields, meshes, lib=lib)
    metricsCalculator.init()
    thermalFields = M.ThermalFields("therm")
    tmodel = M.ThermalModelA(geomFields, thermalFields, meshes, lib=lib)
    tSolver = M.AMG()
    tSolver.relativeTolerance = tol
    tSolver.nMaxIterations = 2000
    tSolver.maxCoarseLevels = 20
    tSolver.verbosity = 2
    if solver in ("bcgstab", "cg"):
        pc = tSolver
        pc.verbosity = 0
        tSolver = M.BCGStab() if solver == "bcgstab" else M.CG()
        tSolver.preconditioner = pc
        tSolver.relativeTolerance = tol
        tSolver.nMaxIterations = 200
    elif solver == "jacobi":
        tSolver = M.JacobiSolver()
        tSolver.relativeTolerance = tol
        tSolver.nMaxIterations = 100000
        tSolver.verbosity = 0
    tmodel.getOptions().linearSolver = tSolver
    return meshes, geomFields, thermalFields, tmodel, tSolver


def test_thermal_script_quad32(devlib, ref, tmp_path):
    raw = G.quad_mesh(32, 32)
    meshes, geomFields, thermalFields, tmodel, tSolver = build_model(devlib, raw, tol=1e-13)
    bcMap = tmodel.getBCMap()
    bc = bcMap[4]
    bc.bcType = "SpecifiedTemperature"
    bc.setVar("specifiedTemperature", 400)
    for gid in (1, 2, 3):
        bcMap[gid].bcType = "SpecifiedTemperature"
        bcMap[gid].setVar("specifiedTemperature", 0)
    for vc in tmodel.getVCMap().values():
        vc.setVar("thermalConductivity", 1.0)
    tmodel.init()
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        tmodel.dumpMatrix("matrix")
        out = io.StringIO()
        with contextlib.redirect_stdout(out):
            tmodel.advance(1)
    finally:
        os.chdir(cwd)
    lines = out.getvalue().splitlines()
    assert lines[0] == "0: [therm.temperature : 63200]"          # same print format as the reference
    assert lines[-1] == "0: [therm.temperature : 63200]"         # ThermalModel's own line
    # reference on the identical mesh
    rm = ref.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes,
                              raw.face_node_count, raw.face_group_size)
    t = ref.RefThermal(rm)
    t.set_bc(4, "SpecifiedTemperature", specifiedTemperature=400)
    for gid in (1, 2, 3):
        t.set_bc(gid, "SpecifiedTemperature", specifiedTemperature=0)
    t.set_solver(ref.solver_cfg(relativeTolerance=1e-13, nMaxIterations=2000, verbosity=0))
    t.init()
    a = t.assemble(1)
    t.advance(1)
    cells = meshes[0].getCells()
    assert rel_l2(thermalFields.temperature[cells], t.field("temperature")) <= 1e-8
    # dumpMatrix writes the reference's file format
    rhs = np.loadtxt(tmp_path / "matrix.rhs")
    assert np.abs(rhs + a["b"][: raw.n_cells]).max() < 1e-6 * np.abs(rhs).max()
    mat = np.loadtxt(tmp_path / "matrix_mesh0.mat", skiprows=2)
    assert mat.shape[1] == 3 and mat[0, 2] == -6.0
    for fg in meshes[0].getBoundaryFaceGroups():
        hf = t.heat_flux(fg.id, fg.site.getCount())
        assert abs(tmodel.getHeatFluxIntegral(meshes[0], fg.id) - hf.sum()) <= 1e-6 * max(abs(hf.sum()), 1.0)


@pytest.mark.parametrize("solver", ["amg", "bcgstab", "cg", "jacobi"])
def test_default_outer_loop_converges(devlib, solver):
    """Default ThermalModel options: outer iterations until ratio < 1e-8 (F/ThermalBC.h:42-43)."""
    raw = G.hex_mesh(10, 10, 10, jitter=0.1, seed=2)
    meshes, geomFields, thermalFields, tmodel, tSolver = build_model(devlib, raw, tol=1e-2, solver=solver)
    tmodel.getBCMap()[5].bcType = "SpecifiedTemperature"
    tmodel.getBCMap()[5]["specifiedTemperature"] = 300.0
    tmodel.getBCMap()[6].bcType = "SpecifiedTemperature"
    tmodel.getBCMap()[6]["specifiedTemperature"] = 400.0
    tmodel.init()
    with contextlib.redirect_stdout(io.StringIO()):
        tmodel.advance(50)
    norms = [t["rnorm"] for t in tmodel.timings]
    assert 2 <= len(norms) < 50 and norms[-1] / norms[0] < 1e-8
    x = thermalFields.temperature[meshes[0].getCells()]
    assert 300.0 - 1e-6 <= x.min() and x.max() <= 400.0 + 1e-6      # discrete maximum principle
    with pytest.raises(M.CException):
        tmodel.getBCMap()[5].setVar("nonsense", 1.0)
    with pytest.raises(M.CException):
        tmodel.getHeatFluxIntegral(meshes[0], 99)
