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
  (void)faceCentroid;  // only the immersed-boundary branches read it (out of scope)
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
