"""Host-side mesh partitioner: one part per GPU, the reference partitioner's numbering semantics.

The reference builds per-rank meshes in `MeshPartitioner` (P/MeshPartitioner.cpp:97-138) from a
ParMETIS cell partition (`ParMETIS_V3_PartMeshKway`, :577-579; ParMETIS 3.1.1 is a missing blob in
the reference tree). ParMETIS only decides WHICH cell goes to WHICH part; everything the
hot path depends on is the local numbering built from that assignment, which is restated here:

  * interior cells of a part in ascending global id              (P/MeshPartitioner.cpp:1606-1635,1665-1836)
  * then one ghost cell per physical-boundary face, grouped by boundary id   (:1749-1779)
  * then one ghost cell per interface face, grouped by neighbour part id     (:1784-1825)
  * faces: interior first, then the boundary groups, then one "interface" group per neighbour
    (F/Mesh.cpp:267-274); at an interface face the ghost is always c1          (:1794-1801)
  * per neighbour: scatter map = local interior cells to send, gather map = local ghost cells
    to fill, both in the SAME (global face id) order on the two sides          (:2019-2102)

The cell assignment itself is "parity unpinned" (any valid partition is acceptable):
`assign_slabs` (structured slabs along the slowest index) and `assign_rcb` (recursive coordinate
bisection, the stand-in for ParMETIS on unstructured meshes) are provided. Solution parity is
checked against the single-partition reference solution (tests/test_partition.py, tests/test_multirank.py).

`hex_slab` builds the local mesh of one slab of a uniform hex box DIRECTLY (no global mesh), for
512^3-class benchmarks; tests check it against `partition_mesh` of the global mesh.
"""
import numpy as np

from . import meshgen
from .meshgen import RawMesh


# ----------------------------------------------------------------------------- cell assignment
def assign_slabs(n_cells, nparts):
    """Contiguous blocks of the global cell numbering (z-slabs for meshgen.hex_mesh)."""
    bounds = (np.arange(nparts + 1, dtype=np.int64) * n_cells) // nparts
    part = np.zeros(n_cells, np.int32)
    for p in range(nparts):
        part[bounds[p]:bounds[p + 1]] = p
    return part


def assign_rcb(centroids, nparts):
    """Recursive coordinate bisection of the cell centroids into `nparts` (any integer) parts."""
    centroids = np.asarray(centroids, np.float64)
    part = np.zeros(len(centroids), np.int32)

    def split(idx, first, count):
        if count == 1:
            part[idx] = first
            return
        left = count // 2
        c = centroids[idx]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        order = np.argsort(c[:, axis], kind="stable")
        k = (len(idx) * left) // count
        split(idx[order[:k]], first, left)
        split(idx[order[k:]], first + left, count - left)

    split(np.arange(len(centroids)), 0, nparts)
    return part


# ----------------------------------------------------------------------------- local mesh of one part
def partition_mesh(raw, geo, part, rank, group_types=None):
    """Local mesh of part `rank`.

    raw  : global RawMesh (meshgen), geo: meshgen.metrics(raw), part: int array [n_cells].
    Returns a RawMesh with the usual connectivity fields plus
      geometry          dict like meshgen.metrics() for the LOCAL numbering (interface ghosts carry
                        the remote cell's centroid and volume)
      group_types       list of group type strings ("interior", "wall"/..., "interface")
      halo              dict(peers, scatter_off, scatter_idx, gather_off, gather_idx)
      cell_global       [n_total] global id of every local cell (boundary ghosts: global ghost id)
      face_global       [n_faces] global face id,  face_flipped [n_faces] bool
    """
    part = np.asarray(part, np.int32)
    fc = raw.face_cells
    nint_g = int(raw.face_group_size[0])
    n_cells_g = raw.n_cells
    # part of both sides of every global face (boundary faces: the ghost belongs to c0's part)
    p0 = part[fc[:, 0]]
    p1 = np.where(fc[:, 1] < n_cells_g, part[np.minimum(fc[:, 1], n_cells_g - 1)], p0)
    mine0, mine1 = p0 == rank, p1 == rank
    faces_g = np.arange(raw.n_faces)
    interior = faces_g[:nint_g][mine0[:nint_g] & mine1[:nint_g]]
    # interface: exactly one side mine, among global interior faces
    xs = faces_g[:nint_g][mine0[:nint_g] ^ mine1[:nint_g]]
    other = np.where(mine0[xs], p1[xs], p0[xs])
    # boundary groups
    b_faces, b_sizes, b_ids, b_types = [], [], [], []
    for g in range(1, len(raw.group_offset)):
        o, c = int(raw.group_offset[g]), int(raw.group_count[g])
        sel = faces_g[o:o + c][mine0[o:o + c]]
        if len(sel):
            b_faces.append(sel); b_sizes.append(len(sel)); b_ids.append(int(raw.group_id[g]))
            b_types.append(group_types[g] if group_types else "wall")
    peers = np.unique(other)
    i_faces = [xs[other == q] for q in peers]          # global face order inside each group
    local_faces = np.concatenate([interior] + b_faces + i_faces).astype(np.int64)
    flipped = np.zeros(len(local_faces), bool)
    n_if = sum(len(f) for f in i_faces)
    if n_if:
        tail = local_faces[len(local_faces) - n_if:]
        flipped[len(local_faces) - n_if:] = ~mine0[tail]   # my cell is the global c1: flip so the ghost is c1
    # local cell numbering
    own = np.nonzero(part == rank)[0]
    n_self = len(own)
    g2l = np.full(n_cells_g, -1, np.int64)
    g2l[own] = np.arange(n_self)
    nb = sum(b_sizes)
    n_total = n_self + nb + n_if
    lfc = np.zeros((len(local_faces), 2), np.int64)
    gf = fc[local_faces]
    c0g = np.where(flipped, gf[:, 1], gf[:, 0])
    c1g = np.where(flipped, gf[:, 0], gf[:, 1])
    lfc[:, 0] = g2l[c0g]
    ni = len(interior)
    lfc[:ni, 1] = g2l[c1g[:ni]]
    lfc[ni:, 1] = n_self + np.arange(nb + n_if)
    cell_global = np.concatenate([own, c1g[ni:]]).astype(np.int64)

    m = RawMesh()
    m.dim = raw.dim
    m.n_cells = int(n_self)
    m.n_total = int(n_total)
    m.n_faces = len(local_faces)
    m.face_cells = np.ascontiguousarray(lfc, np.int32)
    sizes = [ni] + b_sizes + [len(f) for f in i_faces]
    m.face_group_size = np.array(sizes, np.int32)
    m.group_offset = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    m.group_count = m.face_group_size.copy()
    # interface groups get ids after the largest boundary id (the reference numbers them by neighbour)
    base_id = (max(int(i) for i in raw.group_id) + 1) if len(raw.group_id) else 1
    m.group_id = np.array([0] + b_ids + [base_id + int(q) for q in peers], np.int32)
    m.group_types = ["interior"] + b_types + ["interface"] * len(peers)
    m.group_kind = np.array([0] + [3 if t == "symmetry" else 1 for t in b_types] + [2] * len(peers), np.int32)
    m.cell_global = cell_global
    m.face_global = local_faces
    m.face_flipped = flipped
    m.rank, m.nparts = int(rank), int(part.max()) + 1 if len(part) else 1

    # geometry in local numbering
    area = geo["face_area"][local_faces].copy()
    area[flipped] *= -1.0
    ccen = np.zeros((n_total, 3))
    vol = np.zeros(n_total)
    ccen[:n_self] = geo["cell_centroid"][own]
    vol[:n_self] = geo["cell_volume"][own]
    ghosts_g = c1g[ni:]
    ccen[n_self:] = geo["cell_centroid"][ghosts_g]   # boundary ghost: face centroid; interface: remote cell
    vol[n_self:] = geo["cell_volume"][ghosts_g]
    m.geometry = dict(face_area=area, face_area_mag=geo["face_area_mag"][local_faces].copy(),
                      face_centroid=geo["face_centroid"][local_faces].copy(), cell_centroid=ccen,
                      cell_volume=vol)
    # halo maps
    s_off, g_off, s_idx, g_idx = [0], [0], [], []
    pos = n_self + nb
    fpos = ni + nb
    for f in i_faces:
        k = len(f)
        s_idx.append(lfc[fpos:fpos + k, 0])
        g_idx.append(np.arange(pos, pos + k))
        s_off.append(s_off[-1] + k); g_off.append(g_off[-1] + k)
        pos += k; fpos += k
    m.halo = dict(peers=np.array([int(q) for q in peers], np.int32),
                  scatter_off=np.array(s_off, np.int32), gather_off=np.array(g_off, np.int32),
                  scatter_idx=(np.concatenate(s_idx) if s_idx else np.zeros(0)).astype(np.int32),
                  gather_idx=(np.concatenate(g_idx) if g_idx else np.zeros(0)).astype(np.int32))
    return m


def hex_slab(nx, ny, nz, rank, nparts, lx=1.0, ly=1.0, lz=1.0, lib=None):
    """Local mesh of z-slab `rank` of the uniform nx x ny x nz hex box, built without the global
    mesh. Identical (arrays and numbering) to partition_mesh(hex_mesh(nx,ny,nz), assign_slabs).
    With `lib` (a loaded libfvmgpu) the slab's metrics are computed on the device
    (fvmgpu_mesh_compute_geometry) instead of by the numpy restatement."""
    bounds = (np.arange(nparts + 1, dtype=np.int64) * (nx * ny * nz)) // nparts
    if np.any(bounds % (nx * ny)):
        raise ValueError("hex_slab: nz*%d must split into whole z-layers per part" % nparts)
    k0, k1 = int(bounds[rank] // (nx * ny)), int(bounds[rank + 1] // (nx * ny))
    nzl = k1 - k0
    hz = lz / nz
    loc = meshgen.hex_mesh(nx, ny, nzl, lx, ly, hz * nzl)
    loc.nodes[:, 2] += k0 * hz
    if lib is None:
        geo = meshgen.metrics(loc)
    else:
        from . import capi
        row, col = meshgen.connectivity(loc)
        tmp = capi.DeviceMesh(lib, 3, loc.n_cells, loc.n_total, loc.face_cells, row, col, loc.group_offset,
                              loc.group_count, loc.group_id, loc.group_kind)
        geo = tmp.compute_geometry(loc.nodes, loc.face_node_count, loc.face_nodes)
        tmp.close()
    lower, upper = rank > 0, rank < nparts - 1
    # hex_mesh groups: 0 interior, 1..4 sides, 5 z-, 6 z+ ; reference order: boundaries, then interfaces by peer id
    order = [0, 1, 2, 3, 4] + ([] if lower else [5]) + ([] if upper else [6]) + ([5] if lower else []) + ([6] if upper else [])
    types = {5: "interface" if lower else "wall", 6: "interface" if upper else "wall"}
    peers_of = {5: rank - 1, 6: rank + 1}
    face_perm = np.concatenate([np.arange(loc.group_offset[g], loc.group_offset[g] + loc.group_count[g]) for g in order])
    nint = int(loc.group_count[0])
    n_self = loc.n_cells
    fc = loc.face_cells[face_perm].copy()
    nbf = len(face_perm) - nint
    old_ghost = fc[nint:, 1].copy()
    fc[nint:, 1] = n_self + np.arange(nbf)
    m = RawMesh()
    m.dim, m.n_cells, m.n_total, m.n_faces = 3, n_self, n_self + nbf, len(face_perm)
    m.face_cells = np.ascontiguousarray(fc, np.int32)
    sizes = [int(loc.group_count[g]) for g in order]
    m.face_group_size = np.array(sizes, np.int32)
    m.group_offset = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    m.group_count = m.face_group_size.copy()
    gtypes = ["interior"] + [types.get(g, "wall") for g in order[1:]]
    m.group_types = gtypes
    m.group_id = np.array([0] + [(7 + peers_of[g]) if types.get(g) == "interface" else g for g in order[1:]], np.int32)
    m.group_kind = np.array([0] + [2 if t == "interface" else 1 for t in gtypes[1:]], np.int32)
    m.rank, m.nparts = int(rank), int(nparts)
    ccen = np.zeros((m.n_total, 3))
    vol = np.zeros(m.n_total)
    ccen[:n_self] = geo["cell_centroid"][:n_self]
    vol[:n_self] = geo["cell_volume"][:n_self]
    ccen[n_self:] = geo["cell_centroid"][old_ghost]
    vol[n_self:] = geo["cell_volume"][old_ghost]
    s_off, g_off, s_idx, g_idx, peers = [0], [0], [], [], []
    for gi, g in enumerate(order):
        if gi == 0 or types.get(g) != "interface":
            continue
        o, c = int(m.group_offset[gi]), int(m.group_count[gi])
        gh = fc[o:o + c, 1]
        ccen[gh] = ccen[fc[o:o + c, 0]] + np.array([0.0, 0.0, hz if g == 6 else -hz])  # the remote cell
        s_idx.append(fc[o:o + c, 0]); g_idx.append(gh)
        s_off.append(s_off[-1] + c); g_off.append(g_off[-1] + c)
        peers.append(peers_of[g])
    m.geometry = dict(face_area=geo["face_area"][face_perm].copy(), face_area_mag=geo["face_area_mag"][face_perm].copy(),
                      face_centroid=geo["face_centroid"][face_perm].copy(), cell_centroid=ccen, cell_volume=vol)
    m.halo = dict(peers=np.array(peers, np.int32), scatter_off=np.array(s_off, np.int32),
                  gather_off=np.array(g_off, np.int32),
                  scatter_idx=(np.concatenate(s_idx) if s_idx else np.zeros(0)).astype(np.int32),
                  gather_idx=(np.concatenate(g_idx) if g_idx else np.zeros(0)).astype(np.int32))
    own0 = int(bounds[rank])
    m.cell_global = np.concatenate([own0 + np.arange(n_self), np.full(nbf, -1)]).astype(np.int64)
    return m


def block_dims(nparts):
    """(px, py, pz) with px * py * pz = nparts: prime factors dealt out to z, y, x in turn, largest first
    (8 -> 2 x 2 x 2, 4 -> 1 x 2 x 2, 2 -> 1 x 1 x 2) -- what coordinate bisection gives on a uniform box."""
    dims = [1, 1, 1]
    f, n, fac = 2, int(nparts), []
    while n > 1:
        while n % f == 0:
            fac.append(f); n //= f
        f += 1
    for k, q in enumerate(sorted(fac, reverse=True)):
        dims[2 - k % 3] *= q
    return tuple(dims)


def assign_blocks(nx, ny, nz, nparts, cells_per_hex=1):
    """Cell -> part for a lattice mesh of meshgen (hex_mesh: 1 cell per hex, tet_mesh: 6): part = block of the hex."""
    px, py, pz = block_dims(nparts)
    xb, yb, zb = [(np.arange(q + 1, dtype=np.int64) * n) // q for q, n in ((px, nx), (py, ny), (pz, nz))]
    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    b = (np.searchsorted(xb, I.ravel(), side="right") - 1) + px * ((np.searchsorted(yb, J.ravel(), side="right") - 1)
                                                                  + py * (np.searchsorted(zb, K.ravel(), side="right") - 1))
    return np.repeat(b.astype(np.int32), cells_per_hex)


def tet_block(nx, ny, nz, rank, nparts, lx=1.0, ly=1.0, lz=1.0, jitter=0.2, seed=42):
    """Local mesh of block `rank` of meshgen.tet_mesh(nx, ny, nz, ...) cut into block_dims(nparts) blocks, built
    from the block and one layer of hexes around it only -- never the global mesh (a 50 M-tet box is 8 x 6.3 M).
    Identical (arrays, numbering, halo maps, geometry) to partition_mesh(tet_mesh(...), metrics, assign_blocks(...,
    6), rank): global cell ids are lattice functions (6 * hex + tet), interior faces of the generator are ordered
    by their (lower, higher) cell id, and a sub-box keeps that order -- so both sides of an interface list its
    faces in the same order without ever seeing each other's mesh."""
    px, py, pz = block_dims(nparts)
    bx, by, bz = rank % px, (rank // px) % py, rank // (px * py)
    xb, yb, zb = [(np.arange(q + 1, dtype=np.int64) * n) // q for q, n in ((px, nx), (py, ny), (pz, nz))]
    lo = [max(int(xb[bx]) - 1, 0), max(int(yb[by]) - 1, 0), max(int(zb[bz]) - 1, 0)]
    hi = [min(int(xb[bx + 1]) + 1, nx), min(int(yb[by + 1]) + 1, ny), min(int(zb[bz + 1]) + 1, nz)]
    ex, ey, ez = hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]
    hx, hy, hz = lx / nx, ly / ny, lz / nz
    ext = meshgen.tet_mesh(ex, ey, ez, lx=ex * hx, ly=ey * hy, lz=ez * hz, jitter=0.0)
    # node coordinates of the GLOBAL mesh (same linspace values, same jitter stream)
    xs, ys, zs = np.linspace(0.0, lx, nx + 1), np.linspace(0.0, ly, ny + 1), np.linspace(0.0, lz, nz + 1)
    KK, JJ, II = np.meshgrid(np.arange(ez + 1) + lo[2], np.arange(ey + 1) + lo[1], np.arange(ex + 1) + lo[0], indexing="ij")
    II, JJ, KK = II.ravel(), JJ.ravel(), KK.ravel()
    nodes = np.stack([xs[II], ys[JJ], zs[KK]], axis=1)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter, jitter, size=((nx + 1) * (ny + 1) * (nz + 1), 3)) * [hx, hy, hz]
        inner = (II > 0) & (II < nx) & (JJ > 0) & (JJ < ny) & (KK > 0) & (KK < nz)
        gnode = II + (nx + 1) * (JJ + (ny + 1) * KK)
        nodes[inner] += d[gnode[inner]]
        del d
    ext.nodes = nodes
    # owner of every cell of the sub-box, global cell ids
    K, J, I = np.meshgrid(np.arange(ez) + lo[2], np.arange(ey) + lo[1], np.arange(ex) + lo[0], indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    owner = ((np.searchsorted(xb, I, side="right") - 1) + px * ((np.searchsorted(yb, J, side="right") - 1)
                                                                + py * (np.searchsorted(zb, K, side="right") - 1)))
    part = np.repeat(owner.astype(np.int32), 6)
    gid = np.repeat(6 * (I + nx * (J + ny * K.astype(np.int64))), 6) + np.tile(np.arange(6), len(I))
    geo = meshgen.metrics(ext)
    loc = partition_mesh(ext, geo, part, rank)
    cg = loc.cell_global
    loc.cell_global = np.where(cg < ext.n_cells, gid[np.minimum(cg, ext.n_cells - 1)], -1).astype(np.int64)
    loc.nparts = int(nparts)
    loc.nodes_ext = None
    return loc


class MeshPartitioner:
    """Mirror of `fvmparallel.MeshPartitioner(meshes, npart, etype)` (P/MeshPartitioner.i:12-27):
    `partition(); mesh(); meshList()` return THIS rank's meshes. The cell assignment is RCB (or
    slabs for `method="slabs"`) instead of ParMETIS (absent from the image and the reference tree)."""

    def __init__(self, meshes, npart, etype=None, rank=0, method="rcb"):
        self._meshes, self._npart, self._rank, self._method = list(meshes), list(npart), int(rank), method
        self._out = None

    def partition(self):
        self._out = []
        for raw, npart in zip(self._meshes, self._npart):
            geo = meshgen.metrics(raw)
            part = (assign_slabs(raw.n_cells, npart) if self._method == "slabs"
                    else assign_rcb(geo["cell_centroid"][:raw.n_cells], npart))
            self._out.append(partition_mesh(raw, geo, part, self._rank))

    def mesh(self):
        if self._out is None:
            self.partition()
        return self._out

    def meshList(self):
        return self.mesh()
