"""VacancyModel: the vacancy-concentration transport equation on the scalar device path against the
reference's own VacancyModel<double> run in place (oracle/_ref): all four boundary kinds, convection with the
per-face outflow rule, rho * specificVaca time derivative (BDF1 and BDF2)."""
import contextlib
import io

import numpy as np
import pytest

from fvm_b200 import meshgen as G, models as M

SOL_TOL = 1e-9     # north_star: 1e-8; observed 1e-12


@pytest.mark.parametrize("mesh_kind,order", [("hex", 0), ("tet", 1), ("hex", 2)])
def test_vacancy_model_matches_the_reference(devlib, ref, mesh_kind, order):
    raw = G.hex_mesh(6, 7, 5, jitter=0.15, seed=8) if mesh_kind == "hex" else G.tet_mesh(4, 4, 5)
    rm = ref.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes, raw.face_node_count,
                              raw.face_group_size)
    rv = ref.RefVacancy(rm)
    mesh = M.Mesh(raw)
    geom = M.GeomFields("geom")
    M.MeshMetricsCalculatorA(geom, [mesh], lib=devlib).init()
    vf = M.VacancyFields("vacancy")
    vm = M.VacancyModelA(geom, vf, [mesh], lib=devlib)
    bcs = vm.getBCMap()
    for gid, (kind, vars_) in {1: ("SpecifiedVacaFlux", {"specifiedVacaFlux": 2.5}), 2: ("Symmetry", {}),
                               3: ("Convective", {"convectiveCoefficient": 3.0, "farFieldConcentration": 280.0}),
                               4: ("Symmetry", {}), 5: ("SpecifiedConcentration", {"specifiedConcentration": 310.0}),
                               6: ("SpecifiedConcentration", {"specifiedConcentration": 290.0})}.items():
        rv.set_bc(gid, kind, **vars_)
        bcs[gid].bcType = kind
        for k, v in vars_.items():
            bcs[gid][k] = v
    for name, v in (("vacancyDiffusioncoefficient", 0.8), ("density", 2.0), ("specificVaca", 1.5)):
        rv.set_vc(name, v)
        vm.getVCMap()[mesh.getID()][name] = v
    o = vm.getOptions()
    if order:
        rv.set_option("transient", 1); rv.set_option("timeDiscretizationOrder", order); rv.set_option("timeStep", 0.02)
        o.transient, o.timeDiscretizationOrder = True, order
        o["timeStep"] = 0.02
    rv.set_solver(ref.solver_cfg(relativeTolerance=1e-13, nMaxIterations=3000, verbosity=0))
    s = M.AMG()
    s.relativeTolerance, s.nMaxIterations, s.verbosity = 1e-13, 3000, 0
    o.linearSolver = s
    rv.init()
    vm.init()
    flux = rm.geometry()["face_area"] @ np.array([-0.3, 0.5, 0.8])
    rv.field("convectionFlux")[:] = flux
    vf.convectionFlux[mesh.getFaces()][:] = flux
    src = np.cos(np.arange(raw.n_total) * 0.1)
    rv.field("source")[:] = src
    vf.source[mesh.getCells()][:] = src
    text = ""
    for step in range(3 if order else 1):
        rv.advance(2)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            vm.advance(2)
        text = text or buf.getvalue()
        if order:
            rv.update_time()
            vm.updateTime()
    n = raw.n_cells
    got, want = np.asarray(vf.concentration[mesh.getCells()]), rv.field("concentration")
    assert np.abs(got[:n] - want[:n]).max() <= SOL_TOL * np.abs(want[:n]).max()
    for gid in (1, 3, 5, 6):
        a, b = vm.getVacaFluxIntegral(mesh, gid), rv.flux_integral(gid)
        assert abs(a - b) <= 1e-8 * max(1.0, abs(b)), (gid, a, b)
    assert text.splitlines()[0].startswith("0: [vacancy.concentration : ")
    with pytest.raises(M.CException):
        vm.computePlasticStrainRate()
    rv.close()
