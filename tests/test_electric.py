"""ElectricModel parity against the reference's own ElectricModel<double> (fixture
tests/golden/electric_box.npz generated from oracle/_ref by tests/golden/make_golden.py):
electrostatics (potential, electric field), electron velocity, drift face flux (including the
reference's group-local indexing quirk, F/ElectricModel_impl.h:1070-1088) and the charge after drift +
time-derivative transport, over two time steps, through the public ElectricModelA API."""
import contextlib
import io

import numpy as np
import pytest

from conftest import load_golden
from fvm_b200 import meshgen as G, models as M


def rel(a, b):
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def build_model(lib, g):
    raw = G.RawMesh()
    raw.dim, raw.n_cells = 3, int(g["n_self"])
    raw.nodes, raw.face_cells = g["nodes"], g["face_cells"]
    raw.face_nodes, raw.face_node_count, raw.face_group_size = g["face_nodes"], g["face_node_count"], g["face_group_size"]
    raw.n_faces, raw.n_total = len(raw.face_cells), int(g["n_total"])
    raw.group_offset, raw.group_count, raw.group_id, raw.group_kind = (g["group_offset"], g["group_count"],
                                                                       g["group_id"], g["group_kind"])
    mesh = M.Mesh(raw)
    geom = M.GeomFields("geom")
    M.MeshMetricsCalculatorA(geom, [mesh], lib=lib).init()
    ef = M.ElectricFields("elec")
    em = M.ElectricModelA(geom, ef, [mesh], lib=lib)
    bc = em.getBCMap()
    bc[5].bcType = "SpecifiedPotential"; bc[5]["specifiedPotential"] = 0.0
    bc[6].bcType = "SpecifiedPotential"; bc[6]["specifiedPotential"] = 100.0
    bc[1].bcType = "Symmetry"
    bc[2].bcType = "Symmetry"
    bc[3].bcType = "SpecifiedPotentialFlux"; bc[3]["specifiedPotentialFlux"] = 1e-3
    bc[4].bcType = "SpecialDielectricBoundary"; bc[4]["specifiedPotential"] = 20.0
    o = em.getOptions()
    o.drift_enable = True
    o["initialTotalCharge"] = 1e18
    o["timeStep"] = 1e-12
    c = em.getConstants()
    c["nTrap"] = 2
    c["electron_mobility"] = 1e-3
    c["electron_saturation_velocity"] = 1e5
    for nm in ("electrostaticsLinearSolver", "chargetransportLinearSolver"):
        s = M.AMG()
        s.relativeTolerance, s.nMaxIterations, s.verbosity = 1e-13, 2000, 0
        setattr(o, nm, s)
    return mesh, ef, em


def test_electrostatics_and_drift_match_reference(devlib):
    g = load_golden("electric_box.npz")
    mesh, ef, em = build_model(devlib, g)
    em.init()
    cells, faces = mesh.getCells(), mesh.getFaces()
    n = cells.getSelfCount()
    ef.charge[cells][:n, 2] = g["charge0"]
    ef.chargeN1[cells][:] = ef.charge[cells]
    for step in range(2):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            em.advance(1)
        ref_lines = str(g["s%d_text" % step]).strip().splitlines()
        ours = buf.getvalue().strip().splitlines()
        # same residual ratios as the reference prints (its parallel build prints the potential line)
        assert abs(float(ours[0].split(":")[-1].strip(" ];")) - float(ref_lines[0].split(":")[-1].strip(" ];"))) < 1e-6
        assert rel(ef.potential[cells], g["s%d_potential" % step]) <= 1e-10
        assert rel(ef.electric_field[cells], g["s%d_electric_field" % step]) <= 1e-10
        assert rel(ef.electron_velocity[cells], g["s%d_electron_velocity" % step]) <= 1e-10
        assert rel(ef.convectionFlux[faces], g["s%d_convectionFlux" % step]) <= 1e-10
        assert rel(ef.charge[cells], g["s%d_charge" % step]) <= 1e-9
        em.updateTime()


def test_unsupported_source_models_are_rejected(devlib):
    g = load_golden("electric_box.npz")
    mesh, ef, em = build_model(devlib, g)
    em.getOptions().tunneling_enable = True
    with pytest.raises(M.CException):
        em.init()


def test_dielectric_interface_group_and_boundary_match_the_reference(devlib, ref):
    """The two pieces of ElectricModel's electrostatics that need a thin dielectric layer, against the reference run
    in place (oracle/_ref): a face group typed "dielectric interface" (F/DiffusionDiscretization.h:97-151: metric
    sign |A| / (|ds| + thickness / 2), harmonic-mean permittivity, no secondary gradient; its ghost cells keep the
    caller's centroid / volume, F/MeshMetricsCalculator_impl.h:208-209,445-446) and the SpecialDielectricBoundary
    condition with per-face potentials (applyDielectricInterfaceBC, F/GenericBCS.h:367-407). Assembled potential system
    (1e-12 relative) and the solved potential / electric field (1e-10)."""
    raw = G.hex_mesh(6, 5, 4, lx=1e-6, ly=1e-6, lz=0.8e-6, jitter=0.15, seed=5)
    rm = ref.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes, raw.face_node_count,
                              raw.face_group_size, symmetry_groups=(1,), dielectric_groups=(6,))
    conn, geo = rm.connectivity(), rm.geometry()
    assert list(conn["group_kind"]) == [0, 3, 1, 1, 1, 1, 4]
    # the ghost cells of the dielectric-interface group: a cell centre 0.3 h behind the face, the neighbour's volume
    o6, c6 = int(conn["group_offset"][6]), int(conn["group_count"][6])
    fc = conn["face_cells"][o6:o6 + c6]
    ghosts = fc[:, 1]
    assert (np.diff(ghosts) == 1).all()
    en = geo["face_area"][o6:o6 + c6] / geo["face_area_mag"][o6:o6 + c6, None]
    cen = geo["face_centroid"][o6:o6 + c6] + 0.3 * 0.2e-6 * en
    vol = geo["cell_volume"][fc[:, 0]]
    rm.set_cell_geometry(int(ghosts[0]), cen, vol)
    geo = rm.geometry()
    assert np.array_equal(geo["cell_centroid"][ghosts], cen)
    tight = dict(relativeTolerance=1e-13, nMaxIterations=2000, verbosity=0)

    def reference_model():
        e = ref.RefElectric(rm)
        e.set_bc(1, "Symmetry")
        e.set_bc(2, "SpecifiedPotentialFlux", specifiedPotentialFlux=2e-3)
        e.set_bc(3, "SpecialDielectricBoundary", specifiedPotential=20.0)
        e.set_bc(4, "SpecifiedPotentialFlux", specifiedPotentialFlux=0.0)
        e.set_bc(5, "SpecifiedPotential", specifiedPotential=0.0)
        e.set_bc(6, "SpecifiedPotential", specifiedPotential=100.0)
        e.set_option("initialTotalCharge", 1e18)
        e.set_option("chargetransport_enable", 0)
        e.set_constant("dielectric_thickness", 1.5e-7)
        e.set_solver(0, ref.solver_cfg(**tight))
        e.set_solver(1, ref.solver_cfg(**tight))
        e.init()
        return e

    e = reference_model()
    sys_ref = e.potential_system()     # (assembling moves the Dirichlet values into the ghost cells: a model of its own)
    e.close()
    e = reference_model()
    e.advance(2)
    pot_ref, ef_ref = e.field("potential").copy(), e.field("electric_field").reshape(-1, 3).copy()
    e.close()

    m = G.RawMesh()
    m.dim, m.n_cells, m.n_total, m.n_faces = 3, rm.n_self, rm.n_total, rm.n_faces
    m.nodes, m.face_cells = raw.nodes, conn["face_cells"]
    m.face_nodes, m.face_node_count, m.face_group_size = raw.face_nodes, raw.face_node_count, raw.face_group_size
    m.group_offset, m.group_count, m.group_id, m.group_kind = (conn["group_offset"], conn["group_count"], conn["group_id"],
                                                               conn["group_kind"])
    m.geometry = dict(face_area=geo["face_area"], face_area_mag=geo["face_area_mag"], face_centroid=geo["face_centroid"],
                      cell_centroid=geo["cell_centroid"], cell_volume=geo["cell_volume"])
    types = ["interior", "symmetry", "wall", "wall", "wall", "wall", "dielectric interface"]
    mesh = M.Mesh(m, group_types=types)
    geom = M.GeomFields("geom")
    M.MeshMetricsCalculatorA(geom, [mesh], lib=devlib).init()
    ef = M.ElectricFields("elec")
    em = M.ElectricModelA(geom, ef, [mesh], lib=devlib)
    bc = em.getBCMap()
    bc[1].bcType = "Symmetry"
    bc[2].bcType = "SpecifiedPotentialFlux"; bc[2]["specifiedPotentialFlux"] = 2e-3
    bc[3].bcType = "SpecialDielectricBoundary"; bc[3]["specifiedPotential"] = 20.0
    bc[4].bcType = "SpecifiedPotentialFlux"; bc[4]["specifiedPotentialFlux"] = 0.0
    bc[5].bcType = "SpecifiedPotential"; bc[5]["specifiedPotential"] = 0.0
    bc[6].bcType = "SpecifiedPotential"; bc[6]["specifiedPotential"] = 100.0
    o = em.getOptions()
    o["initialTotalCharge"] = 1e18
    o.chargetransport_enable = False
    em.getConstants()["dielectric_thickness"] = 1.5e-7
    for nm in ("electrostaticsLinearSolver", "chargetransportLinearSolver"):
        s = M.AMG()
        s.relativeTolerance, s.nMaxIterations, s.verbosity = 1e-13, 2000, 0
        setattr(o, nm, s)
    em.init()
    cells = mesh.getCells()
    with contextlib.redirect_stdout(io.StringIO()):
        em.advance(2)   # the second iteration sees the gradients (non-orthogonal corrections) of the first
    assert rel(ef.potential[cells], pot_ref) <= 1e-10
    keep = np.ones(rm.n_total, bool)
    keep[ghosts] = False      # (the reference leaves the gradient of the dielectric-interface ghost cells unset)
    assert rel(np.asarray(ef.electric_field[cells])[keep], ef_ref[keep]) <= 1e-10
    # the assembled system itself: re-assemble from the initial state
    em.init()
    ls, _ = em._assemble_electrostatics(mesh)
    ls.lib.timer_stop(1)
    d = ls.download()
    for k, kk in (("diag", "diag"), ("offdiag", "offdiag"), ("b", "b")):
        assert rel(d[k], sys_ref[kk]) <= 1e-12, k
