// fvm_b200 / libfvmgpu -- mesh upload, face<->matrix-entry maps and least-squares gradient
// weights. Setup-time code (runs once per mesh), kept on the device so that a 134 M-cell mesh
// never needs a host-side pass.
#include "structs.cuh"

namespace fvmgpu {

// One thread per face: locate the two off-diagonal entries (c0,c1) and (c1,c0) in the
// cellCells pattern. This is the reference's pair->column map (CRConnectivity::getPairToColMapping,
// F/CRConnectivity.cpp:729-792) built in parallel, plus its inverse entryFace[].
struct PairToColKernel {
  const int* faceCells; const int* row; const int* col; int* pairToCol; int* entryFace; int* errFlag;
  FVM_DEV void operator()(long long f) const {
  const int c0 = faceCells[2 * f], c1 = faceCells[2 * f + 1];
  int p01 = -1, p10 = -1;
  for (int k = row[c0]; k < row[c0 + 1]; k++)
    if (col[k] == c1) { p01 = k; break; }
  for (int k = row[c1]; k < row[c1 + 1]; k++)
    if (col[k] == c0) { p10 = k; break; }
  pairToCol[2 * f] = p01;
  pairToCol[2 * f + 1] = p10;
  if (p01 < 0 || p10 < 0) { atomicOr(errFlag, 1); return; }
  // two faces between the same pair of cells would need the reference's += on one entry
  if (atomicCAS(&entryFace[p01], -1, (int)(2 * f)) != -1) atomicOr(errFlag, 2);
  if (atomicCAS(&entryFace[p10], -1, (int)(2 * f + 1)) != -1) atomicOr(errFlag, 2);
  }
};

struct CheckEntriesKernel {
  const int* entryFace; int* errFlag;
  FVM_DEV void operator()(long long k) const { if (entryFace[k] < 0) atomicOr(errFlag, 4); }
};

struct FaceGroupOfKernel {
  int nInterior, nGroups; const int* gOff; const int* gCnt; int* out;
  FVM_DEV void operator()(long long t) const {
    const int f = nInterior + (int)t;
    int g = -1;
    for (int i = 0; i < nGroups; i++)
      if (f >= gOff[i] && f < gOff[i] + gCnt[i]) g = i;
    out[t] = g;
  }
};

Mesh* meshCreate(int dim, int nSelf, int nTotal, int nFaces, const int* faceCells, const int* ccRow,
                 const int* ccCol, int nGroups, const int* gOff, const int* gCnt, const int* gId,
                 const int* gKind) {
  requireReady();
  if (dim != 2 && dim != 3) fail("mesh: dimension must be 2 or 3 (got %d)", dim);
  if (nSelf <= 0 || nTotal < nSelf || nFaces <= 0) fail("mesh: bad sizes");
  if (nGroups < 1 || gKind[0] != FVMGPU_GROUP_INTERIOR || gOff[0] != 0)
    fail("mesh: the first face group must be the interior group at offset 0 (F/Mesh.cpp:176-187)");
  std::unique_ptr<Mesh> m(new Mesh);
  m->dim = dim;
  m->nSelf = nSelf;
  m->nTotal = nTotal;
  m->nFaces = nFaces;
  m->nnz = ccRow[nTotal];
  m->nInteriorFaces = gCnt[0];
  int expect = 0;
  for (int g = 0; g < nGroups; g++) {
    if (gOff[g] != expect) fail("mesh: face groups must be contiguous and ordered");
    expect += gCnt[g];
    m->groups.push_back({gOff[g], gCnt[g], gId[g], gKind[g]});
  }
  if (expect != nFaces) fail("mesh: face groups cover %d of %d faces", expect, nFaces);
  m->faceCells.upload(faceCells, 2 * (size_t)nFaces);
  m->row.upload(ccRow, (size_t)nTotal + 1);
  m->col.upload(ccCol, (size_t)m->nnz);
  m->entryFace.alloc((size_t)m->nnz);
  m->pairToCol.alloc(2 * (size_t)nFaces);
  m->entryFace.fillBytes(0xff);
  DBuf<int> err(1);
  err.zero();
  parallelFor(nFaces, PairToColKernel{m->faceCells.p, m->row.p, m->col.p, m->pairToCol.p, m->entryFace.p, err.p});
  parallelFor(m->nnz, CheckEntriesKernel{m->entryFace.p, err.p});
  int e = 0;
  err.download(&e, 1);
  if (e & 1) fail("mesh: cellCells does not contain every face's cell pair");
  if (e & 2) fail("mesh: two faces join the same pair of cells (unsupported)");
  if (e & 4) fail("mesh: cellCells has entries that belong to no face");
  // group lookup for non-interior faces
  const int nb = nFaces - m->nInteriorFaces;
  if (nb > 0) {
    DBuf<int> dOff, dCnt;
    dOff.upload(gOff, nGroups);
    dCnt.upload(gCnt, nGroups);
    m->faceGroupOf.alloc(nb);
    parallelFor(nb, FaceGroupOfKernel{m->nInteriorFaces, nGroups, dOff.p, dCnt.p, m->faceGroupOf.p});
    streamSync();
  }
  {
    std::vector<int> kinds(gKind, gKind + nGroups);
    m->groupKindDev.upload(kinds.data(), kinds.size());
    streamSync();
  }
  return m.release();
}

// ---------------------------------------------------------------- geometry
struct Pack4Kernel {
  const double* v3; const double* s; double4* out;
  FVM_DEV void operator()(long long i) const { out[i] = make_double4(v3[3 * i], v3[3 * i + 1], v3[3 * i + 2], s[i]); }
};
struct CheckIbKernel {
  const int* ib; int* err;
  FVM_DEV void operator()(long long i) const { if (ib[i] != -1) atomicOr(err, 1); }
};

// Least-squares gradient weights, one thread per interior row, same operation order as
// GradientModel::getLeastSquaresGradientMatrix3D/2D (F/GradientModel.h:126-282 / :284-436):
//   pass 1  unit = +-ds/|ds| per entry, moment sums in entry order
//   pass 2  w = (K * unit) / |ds|      (K = inverse moment matrix)
//   degenerate cells (det <= eps): w = +-0.5 * A / V_cell
// compiled with -fmad=false so no FMA contraction changes the rounding.
struct LsWeightsKernel {
  int dim; const int* row; const int* col; const int* entryFace; const double4* cellGeom;
  const double4* faceGeom; long long nnz; double* w;
  FVM_DEV void operator()(long long i) const {
  const double4 gi = cellGeom[i];
  const int r0 = row[i], r1 = row[i + 1];
  double Ixx = 0, Iyy = 0, Izz = 0, Ixy = 0, Ixz = 0, Iyz = 0;
  for (int k = r0; k < r1; k++) {
    const int side = entryFace[k] & 1;
    const double4 gn = cellGeom[col[k]];
    // ds = x[c1] - x[c0]
    double d0, d1, d2;
    if (side == 0) { d0 = gn.x - gi.x; d1 = gn.y - gi.y; d2 = gn.z - gi.z; }
    else { d0 = gi.x - gn.x; d1 = gi.y - gn.y; d2 = gi.z - gn.z; }
    const double mag = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    double u0 = d0 / mag, u1 = d1 / mag, u2 = d2 / mag;
    if (side) { u0 = (-d0) / mag; u1 = (-d1) / mag; u2 = (-d2) / mag; }
    Ixx += u0 * u0;
    Iyy += u1 * u1;
    Ixy += u0 * u1;
    if (dim == 3) {
      Izz += u2 * u2;
      Ixz += u0 * u2;
      Iyz += u1 * u2;
    }
  }
  double Kxx = 0, Kxy = 0, Kxz = 0, Kyy = 0, Kyz = 0, Kzz = 0;
  bool degenerate;
  if (dim == 3) {
    const double det = Ixx * (Iyy * Izz - Iyz * Iyz) - Ixy * (Ixy * Izz - Iyz * Ixz) + Ixz * (Ixy * Iyz - Iyy * Ixz);
    degenerate = !(det > 1e-6);
    if (!degenerate) {
      Kxx = (Iyy * Izz - Iyz * Iyz) / det;
      Kxy = -(Ixy * Izz - Iyz * Ixz) / det;
      Kxz = (Ixy * Iyz - Iyy * Ixz) / det;
      Kyy = (Ixx * Izz - Ixz * Ixz) / det;
      Kyz = -(Ixx * Iyz - Ixy * Ixz) / det;
      Kzz = (Ixx * Iyy - Ixy * Ixy) / det;
    }
  } else {
    const double det = Ixx * Iyy - Ixy * Ixy;
    degenerate = !(det > 1e-26);
    if (!degenerate) {
      Kxx = Iyy / det;
      Kxy = -Ixy / det;
      Kyy = Ixx / det;
    }
  }
  for (int k = r0; k < r1; k++) {
    const int ef = entryFace[k];
    const int side = ef & 1;
    double w0, w1, w2;
    if (degenerate) {
      const double4 fg = faceGeom[ef >> 1];
      const double h = side ? -0.5 : 0.5;
      w0 = h * fg.x / gi.w;
      w1 = h * fg.y / gi.w;
      w2 = h * fg.z / gi.w;
    } else {
      const double4 gn = cellGeom[col[k]];
      double d0, d1, d2;
      if (side == 0) { d0 = gn.x - gi.x; d1 = gn.y - gi.y; d2 = gn.z - gi.z; }
      else { d0 = gi.x - gn.x; d1 = gi.y - gn.y; d2 = gi.z - gn.z; }
      const double mag = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
      double u0 = d0 / mag, u1 = d1 / mag, u2 = d2 / mag;
      if (side) { u0 = (-d0) / mag; u1 = (-d1) / mag; u2 = (-d2) / mag; }
      if (dim == 3) {
        w0 = (Kxx * u0 + Kxy * u1 + Kxz * u2) / mag;
        w1 = (Kxy * u0 + Kyy * u1 + Kyz * u2) / mag;
        w2 = (Kxz * u0 + Kyz * u1 + Kzz * u2) / mag;
      } else {
        w0 = (Kxx * u0 + Kxy * u1) / mag;
        w1 = (Kxy * u0 + Kyy * u1) / mag;
        w2 = 0.0 / mag;
      }
    }
    w[k] = w0;
    w[nnz + k] = w1;
    w[2 * nnz + k] = w2;
  }
  }
};

void meshSetGeometry(Mesh* m, const double* faceArea, const double* faceAreaMag, const double* faceCentroid,
                     const double* cellCentroid, const double* cellVolume, const int* ibType) {
  requireReady();
  // face centroids: only the slip-wall boundary condition reads them, and only on boundary faces
  if (faceCentroid && m->nFaces > m->nInteriorFaces)
    m->bFaceCen.upload(faceCentroid + 3 * (size_t)m->nInteriorFaces, 3 * (size_t)(m->nFaces - m->nInteriorFaces));
  if (ibType) {
    DBuf<int> ib, err(1);
    ib.upload(ibType, m->nTotal);
    err.zero();
    parallelFor(m->nTotal, CheckIbKernel{ib.p, err.p});
    int e = 0;
    err.download(&e, 1);
    if (e) fail("mesh: immersed-boundary cells (ibType != IBTYPE_FLUID) are not supported on this path");
  }
  {
    DBuf<double> v3, s;
    v3.upload(cellCentroid, 3 * (size_t)m->nTotal);
    s.upload(cellVolume, (size_t)m->nTotal);
    m->cellGeom.alloc(m->nTotal);
    parallelFor(m->nTotal, Pack4Kernel{v3.p, s.p, m->cellGeom.p});
    v3.upload(faceArea, 3 * (size_t)m->nFaces);
    s.upload(faceAreaMag, (size_t)m->nFaces);
    m->faceGeom.alloc(m->nFaces);
    parallelFor(m->nFaces, Pack4Kernel{v3.p, s.p, m->faceGeom.p});
    streamSync();
  }
  m->gradW.alloc(3 * (size_t)m->nnz);
  m->gradW.zero();
  parallelFor(m->nSelf, LsWeightsKernel{m->dim, m->row.p, m->col.p, m->entryFace.p, m->cellGeom.p, m->faceGeom.p,
                                        m->nnz, m->gradW.p});
  streamSync();
  m->hasGeometry = true;
}

// ================================================================= MeshMetricsCalculator on the device
// face areas (F/MeshMetricsCalculator_impl.h:238-304), area magnitudes (:373-389),
// face centroids with the non-planar correction (:58-120), cell centroids (:128-236: area-weighted
// face centroids, boundary ghosts = the face centroid, reflected on symmetry groups) and cell
// volumes (:392-460: divergence theorem, boundary ghosts = the neighbour's volume). Face loops that
// scatter into cells are per-cell gathers in ascending face order -- the order in which the
// reference's face loop reaches a cell -- and the file is compiled with -fmad=false, so the results
// are bit-identical to the reference's (tests/test_geometry.py).
struct V3g { double x, y, z; };
FVM_DEV V3g ldn(const double* a, int i) { return V3g{a[3 * (size_t)i], a[3 * (size_t)i + 1], a[3 * (size_t)i + 2]}; }
FVM_DEV V3g sub3(V3g a, V3g b) { return V3g{a.x - b.x, a.y - b.y, a.z - b.z}; }
FVM_DEV V3g add3(V3g a, V3g b) { return V3g{a.x + b.x, a.y + b.y, a.z + b.z}; }
FVM_DEV V3g scl3(double s, V3g a) { return V3g{s * a.x, s * a.y, s * a.z}; }
FVM_DEV V3g cross3(V3g a, V3g b) { return V3g{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
FVM_DEV double dotg(V3g a, V3g b) { double s = 0.0; s += a.x * b.x; s += a.y * b.y; s += a.z * b.z; return s; }

struct FaceMetricsKernel {
  const double* nodes; const int* fnOff; const int* fn; double4* faceGeom; double* fcen;
  FVM_DEV void operator()(long long ff) const {
    const int f = (int)ff;
    const int o = fnOff[f], k = fnOff[f + 1] - o;
    V3g A = {0.0, 0.0, 0.0};
    if (k == 2) {
      const V3g dr = sub3(ldn(nodes, fn[o + 1]), ldn(nodes, fn[o]));
      A = V3g{dr.y, -dr.x, 0.0};
    } else if (k == 3) {
      const V3g p0 = ldn(nodes, fn[o]);
      A = scl3(0.5, cross3(sub3(ldn(nodes, fn[o + 1]), p0), sub3(ldn(nodes, fn[o + 2]), p0)));
    } else if (k == 4) {
      A = scl3(0.5, cross3(sub3(ldn(nodes, fn[o + 2]), ldn(nodes, fn[o])), sub3(ldn(nodes, fn[o + 3]), ldn(nodes, fn[o + 1]))));
    } else {
      for (int nn = 0; nn < k; nn++) {
        const V3g n0 = ldn(nodes, fn[o + nn]), n1 = ldn(nodes, fn[o + (nn + 1) % k]);
        const V3g xm = scl3(0.5, add3(n1, n0)), dr = sub3(n1, n0);
        A.x += xm.y * dr.z; A.y += xm.z * dr.x; A.z += xm.x * dr.y;
      }
    }
    const double mag = sqrt(dotg(A, A));
    V3g c = {0.0, 0.0, 0.0};
    if (k > 0) {
      c = ldn(nodes, fn[o]);
      for (int nn = 1; nn < k; nn++) c = add3(c, ldn(nodes, fn[o + nn]));
      const double kk = (double)k;
      c = V3g{c.x / kk, c.y / kk, c.z / kk};
    }
    if (k > 3) {  // correction for non-planar quads and polygons
      const V3g en = {A.x / mag, A.y / mag, A.z / mag};
      double denom = 0.0;
      V3g cfc = {0.0, 0.0, 0.0};
      const double twoThirds = 2. / 3.;
      for (int nn = 0; nn < k; nn++) {
        const V3g n0 = ldn(nodes, fn[o + nn]), n1 = ldn(nodes, fn[o + (nn + 1) % k]);
        const V3g tri = scl3(0.5, cross3(sub3(n0, c), sub3(n1, c)));
        const double triP = dotg(tri, en);
        const V3g xm = scl3(0.5, add3(n0, n1));
        const V3g t = scl3(twoThirds, sub3(xm, c));
        cfc.x += t.x * triP; cfc.y += t.y * triP; cfc.z += t.z * triP;
        denom += triP;
      }
      c.x += cfc.x / denom; c.y += cfc.y / denom; c.z += cfc.z / denom;
    }
    faceGeom[f] = make_double4(A.x, A.y, A.z, mag);
    fcen[3 * (size_t)f] = c.x; fcen[3 * (size_t)f + 1] = c.y; fcen[3 * (size_t)f + 2] = c.z;
  }
};
struct CellCentroidKernel {
  int nSelf, nInteriorFaces; const int* row; const int* col; const int* entryFace; const int* faceGroupOf; const int* groupKind;
  const double4* faceGeom; const double* fcen; double* ccen; int* err;
  FVM_DEV V3g interior(int c) const {
    V3g s = {0.0, 0.0, 0.0};
    double w = 0.0;
    for (int k = row[c]; k < row[c + 1]; k++) {
      const int f = entryFace[k] >> 1;
      const double am = faceGeom[f].w;
      s.x += fcen[3 * (size_t)f] * am; s.y += fcen[3 * (size_t)f + 1] * am; s.z += fcen[3 * (size_t)f + 2] * am;
      w += am;
    }
    return V3g{s.x / w, s.y / w, s.z / w};
  }
  FVM_DEV void operator()(long long ii) const {
    const int i = (int)ii;
    V3g c;
    if (i < nSelf) {
      c = interior(i);
    } else {
      const int k = row[i];
      const int f = entryFace[k] >> 1;
      const int kind = f >= nInteriorFaces ? groupKind[faceGroupOf[f - nInteriorFaces]] : FVMGPU_GROUP_INTERIOR;
      const V3g fc = ldn(fcen, f);
      if (kind == FVMGPU_GROUP_SYMMETRY) {
        const double4 fg = faceGeom[f];
        const V3g en = {fg.x / fg.w, fg.y / fg.w, fg.z / fg.w};
        const V3g c0 = interior(col[k]);
        const V3g dr0 = sub3(fc, c0);
        const double d = dotg(dr0, en);
        const V3g dr1 = {dr0.x - 2. * d * en.x, dr0.y - 2. * d * en.y, dr0.z - 2. * d * en.z};
        c = V3g{c0.x + dr0.x - dr1.x, c0.y + dr0.y - dr1.y, c0.z + dr0.z - dr1.z};
      } else {
        if (kind != FVMGPU_GROUP_BOUNDARY) atomicOr(err, 1);  // interface ghosts get their geometry from the partitioner
        c = fc;
      }
    }
    ccen[3 * (size_t)i] = c.x; ccen[3 * (size_t)i + 1] = c.y; ccen[3 * (size_t)i + 2] = c.z;
  }
};
struct CellVolumeKernel {
  int nSelf; double dim; const int* row; const int* col; const int* entryFace; const double4* faceGeom; const double* fcen;
  const double* ccen; double* vol;
  FVM_DEV double interior(int c) const {
    const V3g cc = ldn(ccen, c);
    double v = 0.0;
    for (int k = row[c]; k < row[c + 1]; k++) {
      const int ef = entryFace[k];
      const int f = ef >> 1;
      const double4 fg = faceGeom[f];
      const double t = dotg(sub3(ldn(fcen, f), cc), V3g{fg.x, fg.y, fg.z}) / dim;
      if (ef & 1) v -= t; else v += t;
    }
    return v;
  }
  FVM_DEV void operator()(long long ii) const {
    const int i = (int)ii;
    vol[i] = i < nSelf ? interior(i) : interior(col[row[i]]);
  }
};
struct PackCellGeomKernel {
  const double* ccen; const double* vol; double4* out;
  FVM_DEV void operator()(long long i) const { out[i] = make_double4(ccen[3 * i], ccen[3 * i + 1], ccen[3 * i + 2], vol[i]); }
};
struct UnpackFaceGeomKernel {
  const double4* fg; double* area; double* mag;
  FVM_DEV void operator()(long long f) const { const double4 g = fg[f]; area[3 * f] = g.x; area[3 * f + 1] = g.y; area[3 * f + 2] = g.z; mag[f] = g.w; }
};

void meshComputeGeometry(Mesh* m, int nNodes, const double* nodes, const int* faceNodeOffsets, const int* faceNodes,
                         double* faceArea, double* faceAreaMag, double* faceCentroid, double* cellCentroid,
                         double* cellVolume) {
  requireReady();
  const size_t nf = (size_t)m->nFaces, nt = (size_t)m->nTotal;
  DBuf<double> dn, fcen(3 * nf), ccen(3 * nt), vol(nt);
  DBuf<int> dOff, dFn, err(1);
  dn.upload(nodes, 3 * (size_t)nNodes);
  dOff.upload(faceNodeOffsets, nf + 1);
  dFn.upload(faceNodes, (size_t)faceNodeOffsets[nf]);
  err.zero();
  m->faceGeom.alloc(nf);
  m->cellGeom.alloc(nt);
  parallelFor((long long)nf, FaceMetricsKernel{dn.p, dOff.p, dFn.p, m->faceGeom.p, fcen.p});
  parallelFor((long long)nt, CellCentroidKernel{m->nSelf, m->nInteriorFaces, m->row.p, m->col.p, m->entryFace.p,
                                                m->faceGroupOf.p, m->groupKindDev.p, m->faceGeom.p, fcen.p, ccen.p, err.p});
  parallelFor((long long)nt, CellVolumeKernel{m->nSelf, (double)m->dim, m->row.p, m->col.p, m->entryFace.p, m->faceGeom.p,
                                              fcen.p, ccen.p, vol.p});
  parallelFor((long long)nt, PackCellGeomKernel{ccen.p, vol.p, m->cellGeom.p});
  int e = 0;
  err.download(&e, 1);
  if (e) fail("compute_geometry: interface ghost cells take their geometry from the partitioner (use fvmgpu_mesh_set_geometry)");
  m->gradW.alloc(3 * (size_t)m->nnz);
  m->gradW.zero();
  parallelFor(m->nSelf, LsWeightsKernel{m->dim, m->row.p, m->col.p, m->entryFace.p, m->cellGeom.p, m->faceGeom.p,
                                        m->nnz, m->gradW.p});
  m->hasGeometry = true;
  if (faceArea || faceAreaMag) {
    DBuf<double> a3(3 * nf), am(nf);
    parallelFor((long long)nf, UnpackFaceGeomKernel{m->faceGeom.p, a3.p, am.p});
    if (faceArea) a3.download(faceArea, 3 * nf);
    if (faceAreaMag) am.download(faceAreaMag, nf);
  }
  if (m->nFaces > m->nInteriorFaces) {
    const size_t nb3 = 3 * (size_t)(m->nFaces - m->nInteriorFaces);
    m->bFaceCen.alloc(nb3);
    copyD2D(m->bFaceCen.p, fcen.p + 3 * (size_t)m->nInteriorFaces, nb3 * sizeof(double));
  }
  if (faceCentroid) fcen.download(faceCentroid, 3 * nf);
  if (cellCentroid) ccen.download(cellCentroid, 3 * nt);
  if (cellVolume) vol.download(cellVolume, nt);
  streamSync();
}

void meshSetHalo(Mesh* m, int nNeigh, const int* peerRank, const int* scatterOff, const int* scatterIdx,
                 const int* gatherOff, const int* gatherIdx) {
  requireReady();
  std::vector<HaloMsg> msgs;
  for (int p = 0; p < nNeigh; p++) {
    HaloMsg hm;
    hm.rank = peerRank[p];
    hm.sendOff = scatterOff[p]; hm.sendCnt = scatterOff[p + 1] - scatterOff[p];
    hm.recvOff = gatherOff[p]; hm.recvCnt = gatherOff[p + 1] - gatherOff[p];
    msgs.push_back(hm);
  }
  const int ns = nNeigh ? scatterOff[nNeigh] : 0, ng = nNeigh ? gatherOff[nNeigh] : 0;
  m->haloScatterHost.assign(scatterIdx, scatterIdx + ns);
  m->haloGatherHost.assign(gatherIdx, gatherIdx + ng);
  for (int v : m->haloScatterHost) if (v < 0 || v >= m->nSelf) fail("set_halo: scatter index %d is not an interior cell", v);
  for (int v : m->haloGatherHost) if (v < m->nSelf || v >= m->nTotal) fail("set_halo: gather index %d is not a ghost cell", v);
  m->halo.build(msgs, m->haloScatterHost, m->haloGatherHost);
  streamSync();
}

}  // namespace fvmgpu
