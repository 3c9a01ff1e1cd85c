"""Shared bodies of the parity tests (used with the CUDA library under -m gpu and with the
test-only host simulator otherwise)."""
import numpy as np

from fvm_b200 import capi as X


def rel_max(u, v):
    return float(np.abs(np.asarray(u) - np.asarray(v)).max() / max(np.abs(v).max(), 1e-300))


def rel_l2(u, v):
    return float(np.linalg.norm(np.asarray(u) - np.asarray(v)) / max(np.linalg.norm(v), 1e-300))


def device_mesh(lib, g):
    """g: dict with the mesh arrays of a golden file or of oracle.refapi (connectivity+geometry)."""
    dm = X.DeviceMesh(lib, int(g["dim"]), int(g["n_self"]), int(g["n_total"]), g["face_cells"], g["cc_row"],
                      g["cc_col"], g["group_offset"], g["group_count"], g["group_id"], g["group_kind"])
    dm.set_geometry(g["face_area"], g["face_area_mag"], g["cell_centroid"], g["cell_volume"],
                    face_centroid=g["face_centroid"], ib_type=g["ib_type"])
    return dm


def ref_mesh_dict(rm):
    d = dict(dim=rm.dim, n_self=rm.n_self, n_total=rm.n_total)
    d.update(rm.connectivity())
    d.update(rm.geometry())
    return d


THERMAL_BC = {
    "SpecifiedTemperature": lambda v: (X.BC_DIRICHLET, [v.get("specifiedTemperature", 300.0)]),
    "SpecifiedHeatFlux": lambda v: (X.BC_NEUMANN, [v.get("specifiedHeatFlux", 0.0)]),
    "Convective": lambda v: (X.BC_CONVECTIVE, [v["convectiveCoefficient"], v["farFieldTemperature"]]),
    "Radiative": lambda v: (X.BC_RADIATIVE, [v["surfaceEmissivity"], v["farFieldTemperature"]]),
    "Mixed": lambda v: (X.BC_MIXED, [v["convectiveCoefficient"], v["surfaceEmissivity"], v["farFieldTemperature"]]),
}


def apply_bcs(ds, bcs):
    for gid, (typ, vals) in bcs.items():
        kind, params = THERMAL_BC[typ](vals)
        ds.set_bc(int(gid), kind, params)


def csr_matvec(row, col, diag, off, x, n):
    """r = diag*x + offdiag*x on the first n rows (numpy, for residual checks in the tests)."""
    lens = np.diff(row[:n + 1])
    rows = np.repeat(np.arange(n), lens)
    e = slice(0, row[n])
    y = diag[:n] * x[:n]
    np.add.at(y, rows, off[e] * x[col[e]])
    return y
