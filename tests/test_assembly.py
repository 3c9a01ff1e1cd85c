"""Assembly parity: gradient, CRMatrix diag/offdiag, residual b, boundary values -- against the
committed reference fixtures and against the live reference on seeded meshes.
Tolerance (north_star): 1e-12 relative; the kernels replay the reference's operation order without
FMA contraction, so the results are in fact expected to be bit-identical."""
import numpy as np
import pytest

from conftest import load_golden
from fvm_b200 import capi as X, meshgen as G
from helpers import apply_bcs, device_mesh, ref_mesh_dict, rel_max

TOL = 1e-12


def check_system(d, ref_diag, ref_off, ref_b):
    assert rel_max(d["diag"], ref_diag) <= TOL
    assert rel_max(d["offdiag"], ref_off) <= TOL
    assert rel_max(d["b"], ref_b) <= TOL


def test_cav32_golden_matrix(devlib):
    """T/THERMAL_MATRIX: bc 3 = 400 K, 4/5/6 = 0 K, k = 1, T0 = 300."""
    g = load_golden("cav32.npz")
    dm = device_mesh(devlib, g)
    assert np.array_equal(dm.pair_to_col(), g["pair_to_col"])
    ds = X.DeviceSystem(devlib, dm)
    ds.fill_field(X.FIELD_X, 300.0)
    ds.set_bc(3, X.BC_DIRICHLET, [400.0])
    for gid in (4, 5, 6):
        ds.set_bc(gid, X.BC_DIRICHLET, [0.0])
    ds.assemble()
    d = ds.download()
    check_system(d, g["diag"], g["off"], g["b"])
    assert np.array_equal(d["diag"], g["diag"]) and np.array_equal(d["offdiag"], g["off"]) and np.array_equal(d["b"], g["b"])
    assert rel_max(ds.get_field(X.FIELD_X), g["x_after_bc"]) == 0.0
    n = int(g["n_self"])
    assert np.abs(-d["b"][:n] - g["golden_rhs"]).max() == 0.0  # the reference's own golden file
    ds.close(); dm.close()


@pytest.mark.parametrize("stage", [0, 1])
def test_hex_every_bc_kind(devlib, stage):
    """Dirichlet / Neumann / convective / radiative / mixed BCs, random conductivity and source on a
    jittered hex mesh; stage 0 = after linearize, stage 1 = after eliminateBoundaryEquations."""
    g = load_golden("hex_bcs.npz")
    dm = device_mesh(devlib, g)
    ds = X.DeviceSystem(devlib, dm)
    ds.set_field(X.FIELD_X, g["x0"])
    ds.set_field(X.FIELD_DIFFUSIVITY, g["k"])
    ds.set_field(X.FIELD_SOURCE, g["src"])
    ds.set_bc(1, X.BC_DIRICHLET, [400.0])
    ds.set_bc(2, X.BC_NEUMANN, [25.0])
    ds.set_bc(3, X.BC_CONVECTIVE, [3.0, 280.0])
    ds.set_bc(4, X.BC_RADIATIVE, [0.8, 250.0])
    ds.set_bc(5, X.BC_MIXED, [2.0, 0.5, 310.0])
    ds.set_bc(6, X.BC_NEUMANN, [0.0])
    ds.assemble(eliminate_boundary=stage)
    d = ds.download()
    p = "s%d_" % stage
    check_system(d, g[p + "diag"], g[p + "offdiag"], g[p + "b"])
    assert np.array_equal(d["is_boundary"], g[p + "is_boundary"])
    assert rel_max(ds.get_field(X.FIELD_GRADIENT), g[p + "gradient"]) <= TOL
    assert rel_max(ds.get_field(X.FIELD_X), g[p + "x"]) <= TOL
    ds.close(); dm.close()


def test_tet_golden(devlib):
    g = load_golden("tet_solve.npz")
    dm = device_mesh(devlib, g)
    ds = X.DeviceSystem(devlib, dm)
    ds.fill_field(X.FIELD_X, 300.0)
    ds.set_field(X.FIELD_DIFFUSIVITY, g["k"])
    ds.set_bc(5, X.BC_DIRICHLET, [300.0])
    ds.set_bc(6, X.BC_DIRICHLET, [400.0])
    ds.set_bc(1, X.BC_NEUMANN, [7.0])
    for gid in (2, 3, 4):
        ds.set_bc(gid, X.BC_NEUMANN, [0.0])
    ds.assemble()
    check_system(ds.download(), g["diag"], g["off"], g["b"])
    ds.close(); dm.close()


CASES = {
    "quad": lambda: G.quad_mesh(19, 13, jitter=0.2, seed=3),
    "quad_ragged": lambda: G.quad_mesh(1, 7),           # every cell touches two boundaries
    "hex": lambda: G.hex_mesh(9, 7, 8, jitter=0.25, seed=4),
    "hex_single_cell": lambda: G.hex_mesh(1, 1, 1),      # one cell, six boundary faces
    "tet": lambda: G.tet_mesh(4, 3, 5, seed=5),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("convecting", [False, True, "central"])
def test_live_reference_random_fields(devlib, ref, case, convecting):
    """Seeded random x / conductivity / source (and a random convecting face flux, which turns the
    Dirichlet groups into the reference's per-face Dirichlet-or-extrapolation switch) on meshes with
    corner cells, a single cell, ragged strips; compared with the reference on identical arrays.
    "central": ConvectionDiscretization's useCentralDifference branch (F/ConvectionDiscretization.h:119-164,
    including its x[c0] + x[c0] face value) instead of the upwind one."""
    m = CASES[case]()
    rm = ref.RefMesh.from_raw(m.dim, m.n_cells, m.nodes, m.face_cells, m.face_nodes, m.face_node_count,
                              m.face_group_size)
    ng = rm.n_groups - 1
    kinds = ["SpecifiedTemperature", "SpecifiedHeatFlux", "Convective", "SpecifiedTemperature", "Mixed", "Radiative"]
    bcs = {}
    for gid in range(1, ng + 1):
        typ = kinds[(gid - 1) % len(kinds)]
        bcs[gid] = (typ, dict(specifiedTemperature=350.0 + gid, specifiedHeatFlux=3.0 * gid,
                              convectiveCoefficient=1.5, farFieldTemperature=290.0, surfaceEmissivity=0.7))
    rng = np.random.default_rng(17)
    x0 = 300.0 + rng.uniform(-20, 20, size=rm.n_total)
    k = np.exp(rng.normal(size=rm.n_total))
    src = 50.0 * rng.normal(size=rm.n_total)
    flux = rng.normal(size=rm.n_faces) * 0.05 if convecting else None
    t = ref.RefThermal(rm)
    for gid, (typ, vals) in bcs.items():
        keep = {"SpecifiedTemperature": ["specifiedTemperature"], "SpecifiedHeatFlux": ["specifiedHeatFlux"],
                "Convective": ["convectiveCoefficient", "farFieldTemperature"],
                "Radiative": ["surfaceEmissivity", "farFieldTemperature"],
                "Mixed": ["convectiveCoefficient", "surfaceEmissivity", "farFieldTemperature"]}[typ]
        t.set_bc(gid, typ, **{kk: vals[kk] for kk in keep})
    t.set_solver(ref.solver_cfg(verbosity=0))
    if convecting == "central":
        t.set_option("useCentralDifference", 1)
    t.init()
    t.field("conductivity")[:] = k
    t.field("source")[:] = src
    if convecting:
        t.field("convectionFlux")[:] = flux
    dm = device_mesh(devlib, ref_mesh_dict(rm))
    ds = X.DeviceSystem(devlib, dm)
    ds.set_field(X.FIELD_DIFFUSIVITY, k)
    ds.set_field(X.FIELD_SOURCE, src)
    if convecting:
        ds.set_field(X.FIELD_FACE_FLUX, flux)
    apply_bcs(ds, bcs)
    if convecting:
        for gid, (typ, vals) in bcs.items():
            if typ == "SpecifiedTemperature":
                ds.set_bc(gid, X.BC_DIRICHLET_OR_OUTFLOW, [vals["specifiedTemperature"]])
    for stage in (0, 1):
        t.field("temperature")[:] = x0
        a = t.assemble(stage)
        ds.set_field(X.FIELD_X, x0)
        ds.assemble(convection={False: 0, True: 1, "central": 2}[convecting], eliminate_boundary=stage)
        d = ds.download()
        check_system(d, a["diag"], a["offdiag"], a["b"])
        assert np.array_equal(d["is_boundary"], a["is_boundary"])
        assert rel_max(ds.get_field(X.FIELD_X), a["x"]) <= TOL
        assert rel_max(ds.get_field(X.FIELD_GRADIENT), t.field("temperatureGradient").reshape(-1, 3)) <= TOL
    ds.close(); dm.close()


def test_transient_terms(devlib, ref):
    """TimeDerivativeDiscretization first and second order (F/TimeDerivativeDiscretization.h)."""
    m = G.hex_mesh(5, 4, 3, jitter=0.1, seed=8)
    rm = ref.RefMesh.from_raw(m.dim, m.n_cells, m.nodes, m.face_cells, m.face_nodes, m.face_node_count,
                              m.face_group_size)
    rng = np.random.default_rng(2)
    for order in (1, 2):
        t = ref.RefThermal(rm)
        t.set_bc(1, "SpecifiedTemperature", specifiedTemperature=400)
        t.set_option("transient", 1)
        t.set_option("timeDiscretizationOrder", order)
        t.set_option("timeStep", 0.01)
        t.set_vc("density", 2.0)
        t.set_vc("specificHeat", 3.0)
        t.set_solver(ref.solver_cfg(verbosity=0))
        t.init()
        x0 = 300 + rng.uniform(-5, 5, rm.n_total)
        x1 = 300 + rng.uniform(-5, 5, rm.n_total)
        x2 = 300 + rng.uniform(-5, 5, rm.n_total)
        t.field("temperature")[:] = x0
        t.field("temperatureN1")[:] = x1
        if order == 2:
            t.field("temperatureN2")[:] = x2
        a = t.assemble(1)
        dm = device_mesh(devlib, ref_mesh_dict(rm))
        ds = X.DeviceSystem(devlib, dm)
        ds.set_field(X.FIELD_X, x0)
        ds.set_field(X.FIELD_X_N1, x1)
        if order == 2:
            ds.set_field(X.FIELD_X_N2, x2)
        ds.set_field(X.FIELD_DENSITY, np.full(rm.n_total, 6.0))
        ds.set_bc(1, X.BC_DIRICHLET, [400.0])
        for gid in range(2, 7):
            ds.set_bc(gid, X.BC_NEUMANN, [0.0])
        ds.assemble(time_order=order, dt=0.01)
        check_system(ds.download(), a["diag"], a["offdiag"], a["b"])
        ds.close(); dm.close()


def test_bad_inputs_raise(devlib):
    g = load_golden("cav32.npz")
    with pytest.raises(X.FvmGpuError, match="dimension"):
        X.DeviceMesh(devlib, 4, int(g["n_self"]), int(g["n_total"]), g["face_cells"], g["cc_row"], g["cc_col"],
                     g["group_offset"], g["group_count"], g["group_id"], g["group_kind"])
    bad_kind = g["group_kind"].copy(); bad_kind[0] = 1
    with pytest.raises(X.FvmGpuError, match="interior"):
        X.DeviceMesh(devlib, 2, int(g["n_self"]), int(g["n_total"]), g["face_cells"], g["cc_row"], g["cc_col"],
                     g["group_offset"], g["group_count"], g["group_id"], bad_kind)
    dm = X.DeviceMesh(devlib, 2, int(g["n_self"]), int(g["n_total"]), g["face_cells"], g["cc_row"], g["cc_col"],
                      g["group_offset"], g["group_count"], g["group_id"], g["group_kind"])
    with pytest.raises(X.FvmGpuError, match="geometry"):
        X.DeviceSystem(devlib, dm)
    ib = g["ib_type"].copy(); ib[5] = -2
    with pytest.raises(X.FvmGpuError, match="immersed"):
        dm.set_geometry(g["face_area"], g["face_area_mag"], g["cell_centroid"], g["cell_volume"], ib_type=ib)
    dm.set_geometry(g["face_area"], g["face_area_mag"], g["cell_centroid"], g["cell_volume"])
    ds = X.DeviceSystem(devlib, dm)
    with pytest.raises(X.FvmGpuError, match="no boundary face group"):
        ds.set_bc(99, X.BC_DIRICHLET, [1.0])
    with pytest.raises(X.FvmGpuError, match="expects"):
        ds.set_field(X.FIELD_X, np.zeros(3))
    with pytest.raises(X.FvmGpuError, match="FACE_FLUX"):
        ds.assemble(convection=1)
    ds.close(); dm.close()


def test_standalone_gradient_is_exact_for_linear_fields(hostsim_lib):
    """GradientModel::compute on its own (fvmgpu_compute_gradient) and the least-squares weights it uses
    (fvmgpu_mesh_download_gradient_weights, F/GradientModel.h:126-436): the LS gradient reproduces a linear field
    exactly on interior cells of a jittered tet mesh, and sum_k w_k (x_nb - x_c) with the downloaded weights is that
    gradient."""
    raw = G.tet_mesh(5, 4, 6, jitter=0.2, seed=11)
    geo = G.metrics(raw)
    row, col = G.connectivity(raw)
    dm = X.DeviceMesh(hostsim_lib, 3, raw.n_cells, raw.n_total, raw.face_cells, row, col, raw.group_offset,
                      raw.group_count, raw.group_id, raw.group_kind)
    dm.set_geometry(geo["face_area"], geo["face_area_mag"], geo["cell_centroid"], geo["cell_volume"],
                    ib_type=np.full(raw.n_total, -1, np.int32))
    a = np.array([1.5, -0.7, 2.25])
    x = geo["cell_centroid"] @ a + 3.0
    ds = X.DeviceSystem(hostsim_lib, dm)
    ds.set_field(X.FIELD_X, x)
    ds.compute_gradient()
    g = ds.get_field(X.FIELD_GRADIENT).reshape(-1, 3)
    # cells all of whose neighbours are interior cells or boundary ghosts holding the linear field
    assert np.abs(g[:raw.n_cells] - a).max() <= 1e-11
    w = dm.gradient_weights()
    i = raw.n_cells // 2
    acc = np.zeros(3)
    for k in range(row[i], row[i + 1]):
        acc += w[k] * (x[col[k]] - x[i])
    assert np.abs(acc - g[i]).max() <= 1e-13
    ds.close(); dm.close()
