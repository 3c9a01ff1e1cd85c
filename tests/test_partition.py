"""Partitioner (fvm_b200/partition.py): the reference MeshPartitioner's local numbering semantics
(P/MeshPartitioner.cpp:1606-1836, 2019-2102) on synthetic meshes."""
import numpy as np
import pytest

from fvm_b200 import meshgen as G, partition as P


@pytest.mark.parametrize("nparts", [2, 3, 4])
def test_direct_slab_mesh_equals_partition_of_the_global_mesh(nparts):
    nx, ny, nz = 5, 4, 12
    raw = G.hex_mesh(nx, ny, nz)
    geo = G.metrics(raw)
    part = P.assign_slabs(raw.n_cells, nparts)
    for r in range(nparts):
        a, b = P.partition_mesh(raw, geo, part, r), P.hex_slab(nx, ny, nz, r, nparts)
        for k in ("n_cells", "n_total", "n_faces"):
            assert a[k] == b[k]
        for k in ("face_cells", "group_offset", "group_count", "group_kind"):
            assert np.array_equal(a[k], b[k]), k
        assert a.group_types == b.group_types
        for k in a.geometry:
            assert np.allclose(a.geometry[k], b.geometry[k], atol=1e-14), k
        for k in a.halo:
            assert np.array_equal(a.halo[k], b.halo[k]), k


def test_direct_slab_mesh_with_device_geometry(hostsim_lib):
    a, b = P.hex_slab(5, 4, 12, 1, 3), P.hex_slab(5, 4, 12, 1, 3, lib=hostsim_lib)
    for k in a.geometry:
        assert np.allclose(a.geometry[k], b.geometry[k], atol=1e-14), k


def test_numbering_follows_the_reference_partitioner():
    raw = G.tet_mesh(4, 3, 5)
    geo = G.metrics(raw)
    part = P.assign_rcb(geo["cell_centroid"][:raw.n_cells], 3)
    assert np.bincount(part).max() - np.bincount(part).min() <= 1
    ms = [P.partition_mesh(raw, geo, part, r) for r in range(3)]
    assert sum(m.n_cells for m in ms) == raw.n_cells
    for r, m in enumerate(ms):
        own = m.cell_global[:m.n_cells]
        assert np.all(np.diff(own) > 0) and np.all(part[own] == r)        # interior cells ascending global id
        kinds = list(m.group_kind)
        assert kinds == sorted(kinds, key=lambda k: {0: 0, 1: 1, 3: 1, 2: 2}[k])  # interior, boundaries, interfaces
        nint = int(m.group_count[0])
        assert np.all(m.face_cells[:nint] < m.n_cells)                   # interior faces touch own cells only
        assert np.array_equal(m.face_cells[nint:, 1], m.n_cells + np.arange(m.n_faces - nint))  # ghost is c1
        h = m.halo
        assert list(h["peers"]) == sorted(h["peers"])
        for i, q in enumerate(h["peers"]):
            s = h["scatter_idx"][h["scatter_off"][i]:h["scatter_off"][i + 1]]
            hq = ms[q].halo
            j = list(hq["peers"]).index(r)
            g = hq["gather_idx"][hq["gather_off"][j]:hq["gather_off"][j + 1]]
            assert np.array_equal(m.cell_global[s], ms[q].cell_global[g])    # same order on both sides
            assert np.all(s < m.n_cells) and np.all(g >= ms[q].n_cells)
        # interface ghosts carry the remote cell's geometry, boundary ghosts the face centroid
        gi = h["gather_idx"]
        assert np.allclose(m.geometry["cell_centroid"][gi], geo["cell_centroid"][m.cell_global[gi]])
        assert np.allclose(m.geometry["cell_volume"][gi], geo["cell_volume"][m.cell_global[gi]])
    # every global interior face appears once as interior or twice as interface (flipped once)
    cnt = np.zeros(raw.n_faces, int)
    for m in ms:
        np.add.at(cnt, m.face_global, 1)
    nint_g = int(raw.face_group_size[0])
    assert set(np.unique(cnt[:nint_g])) <= {1, 2} and np.all(cnt[nint_g:] == 1)


def test_mesh_partitioner_api():
    raw = G.quad_mesh(8, 6)
    mp = P.MeshPartitioner([raw], [2], rank=1)
    mp.partition()
    m = mp.meshList()[0]
    assert m.rank == 1 and m.n_cells == 24 and list(m.halo["peers"]) == [0]


@pytest.mark.parametrize("dims,nparts", [((6, 5, 7), 8), ((5, 4, 6), 2), ((7, 5, 6), 3)])
def test_tet_block_equals_the_partition_of_the_global_mesh(dims, nparts):
    """partition.tet_block builds one block of the jittered tet box from the block and a layer of hexes around it;
    it must be THE part `partition_mesh` cuts out of the global mesh -- same numbering, same halo maps (both sides
    of an interface list its faces in the same order), bit-identical geometry."""
    from fvm_b200 import meshgen as G, partition as P
    nx, ny, nz = dims
    raw = G.tet_mesh(nx, ny, nz, lx=1.3, ly=0.9, lz=1.1)
    geo = G.metrics(raw)
    part = P.assign_blocks(nx, ny, nz, nparts, 6)
    assert sorted(set(part.tolist())) == list(range(nparts))
    for r in range(nparts):
        a = P.partition_mesh(raw, geo, part, r)
        b = P.tet_block(nx, ny, nz, r, nparts, lx=1.3, ly=0.9, lz=1.1)
        assert (a.n_cells, a.n_total, a.n_faces) == (b.n_cells, b.n_total, b.n_faces)
        assert np.array_equal(a.face_cells, b.face_cells)
        for k in ("group_id", "group_count", "group_kind", "group_offset"):
            assert np.array_equal(a[k], b[k]), k
        for k in a.halo:
            assert np.array_equal(a.halo[k], b.halo[k]), k
        for k in a.geometry:
            assert np.array_equal(a.geometry[k], b.geometry[k]), k
        assert np.array_equal(a.cell_global[:a.n_cells], b.cell_global[:b.n_cells])
        gi = a.halo["gather_idx"]
        assert np.array_equal(a.cell_global[gi], b.cell_global[gi])
