"""MeshMetricsCalculator on the device: face areas / centroids, cell centroids /
volumes from the node coordinates, compared with the reference's own MeshMetricsCalculator output
stored in the golden fixtures (bit-identical: same operation order, no FMA contraction)."""
import numpy as np
import pytest

from conftest import load_golden
from fvm_b200 import capi as X, meshgen as G


def _check(lib, g, nodes, face_node_count, face_nodes, exact=True):
    dm = X.DeviceMesh(lib, int(g["dim"]), int(g["n_self"]), int(g["n_total"]), g["face_cells"], g["cc_row"], g["cc_col"],
                      g["group_offset"], g["group_count"], g["group_id"], g["group_kind"])
    out = dm.compute_geometry(nodes, face_node_count, face_nodes)
    for k in ("face_area", "face_area_mag", "face_centroid", "cell_centroid", "cell_volume"):
        ref = np.asarray(g[k]).reshape(out[k].shape)
        if exact:
            assert np.array_equal(out[k], ref), k
        else:
            assert np.abs(out[k] - ref).max() <= 1e-14 * max(np.abs(ref).max(), 1e-300), k
    return dm


def test_jittered_hex_box_with_non_planar_quads(devlib):
    g = load_golden("electric_box.npz")
    dm = _check(devlib, g, g["nodes"], g["face_node_count"], g["face_nodes"])
    # the geometry is installed in the mesh: a system can be assembled on it right away
    ds = X.DeviceSystem(devlib, dm)
    ds.fill_field(X.FIELD_X, 1.0)
    for gid in range(1, 7):
        ds.set_bc(gid, X.BC_NEUMANN, [0.0])
    ds.assemble()
    assert np.abs(ds.download()["b"][: int(g["n_self"])]).max() < 1e-9   # constant field: zero residual
    ds.close(); dm.close()


def test_jittered_quad_mesh(devlib):
    g = load_golden("flow_cavity.npz")
    raw = G.quad_mesh(12, 10, jitter=0.2, seed=5)   # the mesh the fixture was generated on
    _check(devlib, g, raw.nodes, raw.face_node_count, raw.face_nodes).close()


def test_tet_mesh_matches_host_restatement(devlib):
    raw = G.tet_mesh(4, 5, 3)
    mt = G.metrics(raw)     # numpy restatement (accumulation order differs in the volumes: 1e-14)
    row, col = G.connectivity(raw)
    g = dict(dim=3, n_self=raw.n_cells, n_total=raw.n_total, face_cells=raw.face_cells, cc_row=row, cc_col=col,
             group_offset=raw.group_offset, group_count=raw.group_count, group_id=raw.group_id, group_kind=raw.group_kind, **mt)
    _check(devlib, g, raw.nodes, raw.face_node_count, raw.face_nodes, exact=False).close()


def test_partitioned_meshes_are_rejected(devlib):
    from fvm_b200 import partition as P
    raw = G.hex_mesh(4, 4, 6)
    loc = P.hex_slab(4, 4, 6, 0, 2)
    row, col = G.connectivity(loc)
    dm = X.DeviceMesh(devlib, 3, loc.n_cells, loc.n_total, loc.face_cells, row, col, loc.group_offset, loc.group_count,
                      loc.group_id, loc.group_kind)
    with pytest.raises(X.FvmGpuError):
        dm.compute_geometry(raw.nodes, np.full(loc.n_faces, 4, np.int32), np.zeros(4 * loc.n_faces, np.int32))
    dm.close()
