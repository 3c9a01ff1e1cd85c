"""Synthetic mesh generators + host-side mesh metrics (numpy, vectorised).

Produces exactly the arrays the reference's raw `Mesh` constructor takes (F/Mesh.h:93-99,
F/Mesh.cpp:132-247; conventions in SURVEY.md Appendix B):
  * interior faces first, then one boundary group per side (ids 1..G-1, type "wall");
  * one ghost cell per boundary face, numbered nCells + k in boundary-face order;
  * face node order such that the area vector points from c0 to c1
    (F/MeshMetricsCalculator_impl.h:258-285).
`connectivity()` rebuilds cellCells the way the reference does (cellFaces x faceCells with the
neighbours of a cell in ascending face order, F/Mesh.cpp:479-492, F/CRConnectivity.cpp:73-94,
195-266) and `metrics()` restates MeshMetricsCalculator (face areas :238-304, face centroids
:58-120, cell centroids :128-236, volumes :392-460). This is host-side setup that runs once per
mesh (SURVEY §8 "input producer, stays on host"); the tests check it against the reference.
"""
import numpy as np


class RawMesh(dict):
    """dict with attribute access holding the raw Mesh-constructor arrays."""

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _finish(dim, n_cells, nodes, face_cells, face_nodes, face_node_count, group_sizes):
    m = RawMesh()
    m.dim = dim
    m.n_cells = int(n_cells)
    m.nodes = np.ascontiguousarray(nodes, np.float64)
    m.face_cells = np.ascontiguousarray(face_cells, np.int32)
    m.face_nodes = np.ascontiguousarray(face_nodes, np.int32)
    m.face_node_count = np.ascontiguousarray(face_node_count, np.int32)
    m.face_group_size = np.ascontiguousarray(group_sizes, np.int32)
    m.n_faces = len(m.face_cells)
    nb = m.n_faces - int(group_sizes[0])
    m.n_total = m.n_cells + nb
    off = np.concatenate([[0], np.cumsum(group_sizes)[:-1]]).astype(np.int32)
    m.group_offset = off
    m.group_count = m.face_group_size.copy()
    m.group_id = np.arange(len(group_sizes), dtype=np.int32)  # raw ctor: id = group index
    m.group_kind = np.array([0] + [1] * (len(group_sizes) - 1), np.int32)
    return m


def quad_mesh(nx, ny, lx=1.0, ly=1.0, jitter=0.0, seed=0):
    """nx x ny quadrilaterals on [0,lx]x[0,ly]; boundary groups: 1 left, 2 right, 3 bottom, 4 top."""
    xs = np.linspace(0.0, lx, nx + 1)
    ys = np.linspace(0.0, ly, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # [j, i]
    nodes = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], axis=1)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        hx, hy = lx / nx, ly / ny
        inner = np.zeros((ny + 1, nx + 1), bool)
        inner[1:-1, 1:-1] = True
        d = rng.uniform(-jitter, jitter, size=(nodes.shape[0], 2)) * [hx, hy]
        nodes[inner.ravel(), :2] += d[inner.ravel()]

    def nid(i, j):
        return i + (nx + 1) * j

    def cid(i, j):
        return i + nx * j

    I, J = np.meshgrid(np.arange(nx - 1), np.arange(ny), indexing="xy")
    I, J = I.ravel(), J.ravel()
    fx_cells = np.stack([cid(I, J), cid(I + 1, J)], 1)
    fx_nodes = np.stack([nid(I + 1, J), nid(I + 1, J + 1)], 1)
    I2, J2 = np.meshgrid(np.arange(nx), np.arange(ny - 1), indexing="xy")
    I2, J2 = I2.ravel(), J2.ravel()
    fy_cells = np.stack([cid(I2, J2), cid(I2, J2 + 1)], 1)
    fy_nodes = np.stack([nid(I2 + 1, J2 + 1), nid(I2, J2 + 1)], 1)
    jj = np.arange(ny)
    ii = np.arange(nx)
    bl_c, bl_n = cid(0, jj), np.stack([nid(0, jj + 1), nid(0, jj)], 1)
    br_c, br_n = cid(nx - 1, jj), np.stack([nid(nx, jj), nid(nx, jj + 1)], 1)
    bb_c, bb_n = cid(ii, 0), np.stack([nid(ii, 0), nid(ii + 1, 0)], 1)
    bt_c, bt_n = cid(ii, ny - 1), np.stack([nid(ii + 1, ny), nid(ii, ny)], 1)
    n_cells = nx * ny
    n_int = len(fx_cells) + len(fy_cells)
    bc0 = np.concatenate([bl_c, br_c, bb_c, bt_c])
    ghosts = n_cells + np.arange(len(bc0))
    face_cells = np.concatenate([fx_cells, fy_cells, np.stack([bc0, ghosts], 1)])
    face_nodes = np.concatenate([fx_nodes, fy_nodes, bl_n, br_n, bb_n, bt_n])
    fnc = np.full(len(face_cells), 2, np.int32)
    return _finish(2, n_cells, nodes, face_cells, face_nodes.ravel(), fnc, [n_int, ny, ny, nx, nx])


def hex_mesh(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, jitter=0.0, seed=0):
    """nx x ny x nz hexahedra on a box; boundary groups: 1 x-, 2 x+, 3 y-, 4 y+, 5 z-, 6 z+."""
    xs = np.linspace(0.0, lx, nx + 1)
    ys = np.linspace(0.0, ly, ny + 1)
    zs = np.linspace(0.0, lz, nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")  # [k, j, i]
    nodes = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        inner = np.zeros((nz + 1, ny + 1, nx + 1), bool)
        inner[1:-1, 1:-1, 1:-1] = True
        d = rng.uniform(-jitter, jitter, size=nodes.shape) * [lx / nx, ly / ny, lz / nz]
        nodes[inner.ravel()] += d[inner.ravel()]
    npx, npy = nx + 1, ny + 1

    def nid(i, j, k):
        return (i + npx * (j + npy * k)).astype(np.int64)

    def cid(i, j, k):
        return (i + nx * (j + ny * k)).astype(np.int64)

    def grid(ni, nj, nk):
        K, J, I = np.meshgrid(np.arange(nk), np.arange(nj), np.arange(ni), indexing="ij")
        return I.ravel(), J.ravel(), K.ravel()

    def xface(i, j, k, flip=False):  # quad in the y-z plane at node plane i, normal +x
        n = [nid(i, j, k), nid(i, j + 1, k), nid(i, j + 1, k + 1), nid(i, j, k + 1)]
        if flip:
            n = [n[0], n[3], n[2], n[1]]
        return np.stack(n, 1)

    def yface(i, j, k, flip=False):  # normal +y
        n = [nid(i, j, k), nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j, k)]
        if flip:
            n = [n[0], n[3], n[2], n[1]]
        return np.stack(n, 1)

    def zface(i, j, k, flip=False):  # normal +z
        n = [nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i, j + 1, k)]
        if flip:
            n = [n[0], n[3], n[2], n[1]]
        return np.stack(n, 1)

    cells, fnodes = [], []
    I, J, K = grid(nx - 1, ny, nz)
    cells.append(np.stack([cid(I, J, K), cid(I + 1, J, K)], 1)); fnodes.append(xface(I + 1, J, K))
    I, J, K = grid(nx, ny - 1, nz)
    cells.append(np.stack([cid(I, J, K), cid(I, J + 1, K)], 1)); fnodes.append(yface(I, J + 1, K))
    I, J, K = grid(nx, ny, nz - 1)
    cells.append(np.stack([cid(I, J, K), cid(I, J, K + 1)], 1)); fnodes.append(zface(I, J, K + 1))
    n_int = sum(len(c) for c in cells)
    n_cells = nx * ny * nz
    bcells, bsizes = [], []
    zero = lambda a: np.zeros_like(a)
    I, J, K = grid(1, ny, nz)
    bcells.append(cid(zero(I), J, K)); fnodes.append(xface(zero(I), J, K, flip=True)); bsizes.append(len(J))
    bcells.append(cid(zero(I) + nx - 1, J, K)); fnodes.append(xface(zero(I) + nx, J, K)); bsizes.append(len(J))
    I, J, K = grid(nx, 1, nz)
    bcells.append(cid(I, zero(J), K)); fnodes.append(yface(I, zero(J), K, flip=True)); bsizes.append(len(I))
    bcells.append(cid(I, zero(J) + ny - 1, K)); fnodes.append(yface(I, zero(J) + ny, K)); bsizes.append(len(I))
    I, J, K = grid(nx, ny, 1)
    bcells.append(cid(I, J, zero(K))); fnodes.append(zface(I, J, zero(K), flip=True)); bsizes.append(len(I))
    bcells.append(cid(I, J, zero(K) + nz - 1)); fnodes.append(zface(I, J, zero(K) + nz)); bsizes.append(len(I))
    bc0 = np.concatenate(bcells)
    ghosts = n_cells + np.arange(len(bc0))
    cells.append(np.stack([bc0, ghosts], 1))
    face_cells = np.concatenate(cells)
    face_nodes = np.concatenate(fnodes)
    fnc = np.full(len(face_cells), 4, np.int32)
    return _finish(3, n_cells, nodes, face_cells, face_nodes.ravel(), fnc, [n_int] + bsizes)


def tet_mesh(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, jitter=0.2, seed=42):
    """Each hexahedron of an nx x ny x nz box split into 6 tetrahedra (Kuhn triangulation, conforming
    across neighbours), interior nodes jittered by U(-jitter, jitter) h. Boundary groups 1..6 as
    hex_mesh. Faces are found by sorting node triples."""
    xs = np.linspace(0.0, lx, nx + 1)
    ys = np.linspace(0.0, ly, ny + 1)
    zs = np.linspace(0.0, lz, nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    nodes = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if jitter > 0:
        rng = np.random.default_rng(seed)
        inner = np.zeros((nz + 1, ny + 1, nx + 1), bool)
        inner[1:-1, 1:-1, 1:-1] = True
        d = rng.uniform(-jitter, jitter, size=nodes.shape) * [lx / nx, ly / ny, lz / nz]
        nodes[inner.ravel()] += d[inner.ravel()]
    npx, npy = nx + 1, ny + 1
    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()

    def nid(di, dj, dk):
        return (I + di) + npx * ((J + dj) + npy * (K + dk))

    # Kuhn: the 6 monotone paths from corner (0,0,0) to (1,1,1)
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets = []
    for p in perms:
        v = [np.zeros(3, int)]
        for ax in p:
            e = v[-1].copy()
            e[ax] = 1
            v.append(e)
        tets.append(np.stack([nid(*v[0]), nid(*v[1]), nid(*v[2]), nid(*v[3])], 1))
    tets = np.stack(tets, 1).reshape(-1, 4)  # cell id = 6*hex + t
    n_cells = len(tets)
    # faces: node triples opposite each vertex
    loc = np.array([[1, 2, 3], [0, 3, 2], [0, 1, 3], [0, 2, 1]])
    tri = tets[:, loc]  # [cell, 4, 3]
    cell_of = np.repeat(np.arange(n_cells), 4)
    tri = tri.reshape(-1, 3)
    key = np.sort(tri, axis=1)
    nn = nodes.shape[0]
    k64 = (key[:, 0].astype(np.int64) * nn + key[:, 1]) * nn + key[:, 2]
    order = np.argsort(k64, kind="stable")
    ks = k64[order]
    same_next = np.zeros(len(ks), bool)
    same_next[:-1] = ks[1:] == ks[:-1]
    same_prev = np.zeros(len(ks), bool)
    same_prev[1:] = same_next[:-1]
    first_of_pair = order[same_next]
    second_of_pair = order[np.nonzero(same_next)[0] + 1]
    single = order[~same_next & ~same_prev]
    # interior faces: c0 = lower cell id, nodes as seen from c0 (oriented outward from c0 below)
    a, b = cell_of[first_of_pair], cell_of[second_of_pair]
    swap = a > b
    c0 = np.where(swap, b, a)
    c1 = np.where(swap, a, b)
    src = np.where(swap, second_of_pair, first_of_pair)
    so = np.argsort(c0 * np.int64(n_cells) + c1, kind="stable")
    c0, c1, src = c0[so], c1[so], src[so]
    int_nodes = tri[src]
    # boundary faces grouped by side
    bcell = cell_of[single]
    bn = tri[single]
    cen = nodes[bn].mean(axis=1)
    eps = 1e-9
    side = np.full(len(single), -1)
    for g, (ax, val) in enumerate([(0, 0.0), (0, lx), (1, 0.0), (1, ly), (2, 0.0), (2, lz)]):
        side[np.abs(cen[:, ax] - val) < eps * max(lx, ly, lz)] = g
    if (side < 0).any():
        raise RuntimeError("tet_mesh: unclassified boundary face")
    bo = np.lexsort((bcell, side))
    bcell, bn, side = bcell[bo], bn[bo], side[bo]
    bsizes = [int((side == g).sum()) for g in range(6)]
    ghosts = n_cells + np.arange(len(bcell))
    face_cells = np.concatenate([np.stack([c0, c1], 1), np.stack([bcell, ghosts], 1)])
    face_nodes = np.concatenate([int_nodes, bn])
    # orient: area (n1-n0)x(n2-n0)/2 must point away from c0's centroid
    p = nodes[face_nodes]
    area = 0.5 * np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    cc = nodes[tets].mean(axis=1)
    outward = np.einsum("ij,ij->i", p.mean(axis=1) - cc[face_cells[:, 0]], area) > 0
    flip = ~outward
    face_nodes[flip] = face_nodes[flip][:, [0, 2, 1]]
    fnc = np.full(len(face_cells), 3, np.int32)
    return _finish(3, n_cells, nodes, face_cells, face_nodes.ravel(), fnc, [len(c0)] + bsizes)


def connectivity(m):
    """cellCells CSR (row, col) in the reference's order: neighbours of a cell in ascending face
    order (cellFaces = transpose(faceCells); cellCells = cellFaces x faceCells, diagonal implicit)."""
    fc = m.face_cells
    nt = m.n_total
    cell = fc.reshape(-1)
    other = fc[:, ::-1].reshape(-1)
    order = np.argsort(cell, kind="stable")  # faces already ascending inside each cell's run
    counts = np.bincount(cell, minlength=nt)
    row = np.zeros(nt + 1, np.int32)
    np.cumsum(counts, out=row[1:])
    col = other[order].astype(np.int32)
    return row, col


def metrics(m):
    """Face areas / centroids, cell centroids / volumes (restates MeshMetricsCalculator)."""
    nodes = m.nodes
    fc = m.face_cells
    nf = m.n_faces
    k = int(m.face_node_count[0])
    if not (m.face_node_count == k).all():
        raise NotImplementedError("mixed face types")
    fn = m.face_nodes.reshape(nf, k)
    P = nodes[fn]  # [F, k, 3]
    if k == 2:
        dr = P[:, 1] - P[:, 0]
        area = np.stack([dr[:, 1], -dr[:, 0], np.zeros(nf)], 1)
    elif k == 3:
        area = 0.5 * np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
    elif k == 4:
        area = 0.5 * np.cross(P[:, 2] - P[:, 0], P[:, 3] - P[:, 1])
    else:
        raise NotImplementedError("polygonal faces")
    area_mag = np.sqrt((area * area).sum(1))
    fcen = P[:, 0].copy()
    for j in range(1, k):
        fcen += P[:, j]
    fcen /= float(k)
    if k > 3:  # non-planar quad correction, F/MeshMetricsCalculator_impl.h:88-114
        en = area / area_mag[:, None]
        denom = np.zeros(nf)
        cfc = np.zeros((nf, 3))
        for j in range(k):
            n0, n1 = P[:, j], P[:, (j + 1) % k]
            tri = 0.5 * np.cross(n0 - fcen, n1 - fcen)
            tap = (tri * en).sum(1)
            xm = 0.5 * (n0 + n1)
            cfc += (2.0 / 3.0) * (xm - fcen) * tap[:, None]
            denom += tap
        fcen = fcen + cfc / denom[:, None]
    nt, ns = m.n_total, m.n_cells
    ccen = np.zeros((nt, 3))
    w = np.zeros(nt)
    for s in (0, 1):
        c = fc[:, s]
        for d in range(3):
            ccen[:, d] += np.bincount(c, weights=fcen[:, d] * area_mag, minlength=nt)
        w += np.bincount(c, weights=area_mag, minlength=nt)
    ccen[:ns] /= w[:ns, None]
    nint = int(m.face_group_size[0])
    ccen[fc[nint:, 1]] = fcen[nint:]  # boundary ghost = face centroid (groups are all "wall")
    dim = float(m.dim)
    vol = np.zeros(nt)
    vol += np.bincount(fc[:, 0], weights=((fcen - ccen[fc[:, 0]]) * area).sum(1) / dim, minlength=nt)
    vol -= np.bincount(fc[:, 1], weights=((fcen - ccen[fc[:, 1]]) * area).sum(1) / dim, minlength=nt)
    vol[fc[nint:, 1]] = vol[fc[nint:, 0]]
    return dict(face_area=area, face_area_mag=area_mag, face_centroid=fcen, cell_centroid=ccen,
                cell_volume=vol)
