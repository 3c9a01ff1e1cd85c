// fvm_b200 / libfvmgpu -- FlowModel (SIMPLE) hot path on the device, FP64, sm_100a.
//
// Per outer iteration the reference runs (F/FlowModel_impl.h:1433-1471)
//   solveMomentum   (:730-770)  velocity gradient, Diffusion<Vec3,DiagTensor3,T> + Convection +
//                               MomentumPressureGradient (+TimeDerivative) + BCs + Underrelaxer,
//                               AMG on CRMatrix<DiagTensor3,T,Vec3>, x += delta, momAp = diag
//   solveContinuity (:1410-1430) Rhie-Chow face mass flux + pressure-correction matrix
//                               (F/FlowModelInterior.h:8-218), fixed-flux boundaries
//                               (F/FlowModelVelocityBC.h:11-103), net-flux redistribution and the
//                               reference cell (:1133-1188, :906-994), AMG, then correctPressure,
//                               correctMassFlux, correctVelocity, updateFacePressure,
//                               computeContinuityResidual (:1263-1339)
// Here every face loop that scatters into cells is a row-parallel GATHER over the cell's faces in
// ascending face order (the order in which the reference's scatter reaches that cell), so all sums
// are deterministic; loops whose outputs are per face (mass flux, pCoeff, face pressure) are
// face-parallel. Compiled with -fmad=false like assemble.cu (same IEEE operation sequence as the
// reference's x86-64 build).
//
// Boundary types: NoSlipWall (applyDirichletBC), Symmetry (vector applySymmetryBC + reflected gradients),
// VelocityBoundary and PressureBoundary (per-face extrapolation / Dirichlet, fixedPressureMomentumBC,
// fixedPressureContinuityBC, boundary mass-flux rows, pressureBoundaryPostContinuitySolve); SlipJump,
// turbulence and the PV-coupled solve are not built.
// Several GPUs (one mesh part each): partition-interface faces are treated like interior faces
// everywhere (the reference loops discretizeMassFluxInterior / correctMassFluxInterior /
// correctVelocityInterior / updateFacePressureInterior over mesh.getInterfaceGroups(),
// F/FlowModel_impl.h:1027-1035,1297-1309), the ghost copies of V, p, grad V, grad p and momAp are
// refreshed where the reference calls syncLocal (:768, :1011, :1334-1335) or its gradient / linear-system
// syncs, net flux / volume / norms are all-reduced (:1134-1141, :1167-1169) and the reference pressure
// correction comes from the rank that owns the globally lowest cell (:931-994, :1216-1230).
// The momentum system's diagonal is a DiagonalTensor (one value per velocity component) with a
// shared scalar off-diagonal: the three components are solved one after the other by the scalar
// AMG on ONE hierarchy when their diagonals coincide (always the case without symmetry planes).
#include "solver.cuh"

namespace fvmgpu {

struct FlowBcEntry {
  int offset, count;
  int kind;       // FVMGPU_FLOWBC_* or -1
  int groupKind;
  double p[5];    // vx, vy, vz, specifiedPressure, accomodationCoefficient
};

struct Flow {
  Mesh* mesh = nullptr;
  int nSelf = 0, nTotal = 0, nFaces = 0;
  long long nnz = 0;
  DBuf<double> V, Vprev, VN1, VN2;          // 3*Nt AoS
  DBuf<double> p, pFace, rho, mu, massFlux, contResid;
  DBuf<double> pGrad;                        // 3*Nt
  DBuf<double> vGrad;                        // 9*Nt  [cell][direction i][component k]
  DBuf<double> momAp;                        // 3*Nt
  DBuf<double> mDiag, mOff, mB, mDelta;      // momentum system: diag/b/delta 3*Nt AoS, off nnz
  DBuf<double> pCoeff;                       // per face
  DBuf<double> bDiagAdd, bOff10, bCoeffL;     // per boundary face: pressure-boundary continuity coefficients
  bool hasPressureBoundary = false;
  std::unique_ptr<System> comp;              // scalar system reused for each velocity component
  std::unique_ptr<System> pp;                // pressure-correction system
  DBuf<double> lastDiag;                     // diagonal the current momentum hierarchy was built for
  std::vector<FlowBcEntry> bcs;
  DBuf<FlowBcEntry> bcsDev;
  bool bcsDirty = true;
  bool hasMomAp = false, hasVN1 = false, hasVN2 = false;
  int refCell = 0;                           // local index of the reference cell, -1: another rank owns it
  bool multi = false;                        // mesh part of a partitioned mesh: halo exchanges + all-reduces
  bool pressureBoundaryAnywhere = false;     // hasPressureBoundary or-ed over the ranks (agreed in flowInit)
  DBuf<double> scal;                         // device scalars: [0] netFlux, [1] volumeSum, [2..4] norms, [6] reference pp
};

// ---------------------------------------------------------------- small helpers
struct V3 { double x, y, z; };
FVM_DEV V3 ld3(const double* a, int i) { V3 v; v.x = a[3 * (size_t)i]; v.y = a[3 * (size_t)i + 1]; v.z = a[3 * (size_t)i + 2]; return v; }
FVM_DEV void st3(double* a, int i, const V3& v) { a[3 * (size_t)i] = v.x; a[3 * (size_t)i + 1] = v.y; a[3 * (size_t)i + 2] = v.z; }
FVM_DEV double dot3(const V3& a, const V3& b) {  // Vector::dot: sum over components in order (F/Vector.h)
  double s = 0.0;
  s += a.x * b.x; s += a.y * b.y; s += a.z * b.z;
  return s;
}
FVM_DEV double harmonicAvg(double x0, double x1) {  // F/DiffusionDiscretization.h:19-27
  const double sum = x0 + x1;
  if (x0 + x1 != 0.0) return 2.0 * x0 * x1 / sum;
  return sum;
}

FVM_DEV const FlowBcEntry* faceBc(const FlowBcEntry* bcs, const int* faceGroupOf, int nInteriorFaces, int f) {
  return &bcs[faceGroupOf[f - nInteriorFaces]];
}
// interior faces and partition-interface faces (whose c1 is the ghost copy of a neighbour rank's cell)
FVM_DEV bool interiorLike(const FlowBcEntry* bcs, const int* faceGroupOf, int nInteriorFaces, int f) {
  return f < nInteriorFaces || bcs[faceGroupOf[f - nInteriorFaces]].groupKind == FVMGPU_GROUP_INTERFACE;
}

// ---------------------------------------------------------------- init: default mass flux
struct FlowInitMassFluxFaces {  // FlowModel::init, F/FlowModel_impl.h:222-244, 297-312
  int nInteriorFaces; const int* faceCells; const int* faceGroupOf; const double4* faceGeom; const double* V;
  const double* rho; const FlowBcEntry* bcs; double* massFlux;
  FVM_DEV void operator()(long long ff) const {
    const int f = (int)ff;
    const int c0 = faceCells[2 * f], c1 = faceCells[2 * f + 1];
    const double4 fg = faceGeom[f];
    const V3 A = {fg.x, fg.y, fg.z};
    if (f >= nInteriorFaces) {
      const FlowBcEntry* bc = faceBc(bcs, faceGroupOf, nInteriorFaces, f);
      if (bc->kind == FVMGPU_FLOWBC_NOSLIP_WALL || bc->kind == FVMGPU_FLOWBC_VELOCITY ||
          bc->kind == FVMGPU_FLOWBC_SLIP_JUMP) {  // F/FlowModel_impl.h:297-312
        const V3 bv = {bc->p[0], bc->p[1], bc->p[2]};
        massFlux[f] = rho[c0] * dot3(bv, A);
        return;
      }
      if (bc->kind == FVMGPU_FLOWBC_SYMMETRY) { massFlux[f] = 0.0; return; }  // F/FlowModel_impl.h:324-328
    }
    massFlux[f] = 0.5 * (rho[c0] * dot3(ld3(V, c0), A) + rho[c1] * dot3(ld3(V, c1), A));
  }
};

struct ContResidRows {  // computeContinuityResidual, F/FlowModel_impl.h:1235-1261
  const int* row; const int* entryFace; const double* massFlux; double* r;
  FVM_DEV void operator()(long long ii) const {
    const int i = (int)ii;
    double s = 0.0;
    for (int k = row[i]; k < row[i + 1]; k++) {
      const int ef = entryFace[k];
      const double mf = massFlux[ef >> 1];
      if (ef & 1) s -= mf; else s += mf;
    }
    r[i] = s;
  }
};

// ---------------------------------------------------------------- gradients
// GradientModel<Vector<T,3>>::compute: g[i] += w[i] * (x_nb - x_c)   (F/GradientMatrix.h:55-76)
FVM_DEV void velGradOf(int c, const int* row, const int* col, const double* V, const double* w, long long nnz,
                       double* g /*9*/) {
#pragma unroll
  for (int q = 0; q < 9; q++) g[q] = 0.0;
  const V3 xc = ld3(V, c);
  for (int k = row[c]; k < row[c + 1]; k++) {
    const V3 xn = ld3(V, col[k]);
    const double d0 = xn.x - xc.x, d1 = xn.y - xc.y, d2 = xn.z - xc.z;
    const double w0 = w[k], w1 = w[nnz + k], w2 = w[2 * nnz + k];
    g[0] += w0 * d0; g[1] += w0 * d1; g[2] += w0 * d2;
    g[3] += w1 * d0; g[4] += w1 * d1; g[5] += w1 * d2;
    g[6] += w2 * d0; g[7] += w2 * d1; g[8] += w2 * d2;
  }
}
// ghost cells copy their neighbour's gradient, or reflect it on face groups of kind SYMMETRY
// (F/GradientModel.h:530-566; reflectGradient for Vector<T,3>: GTR = R * GT0 * R, :62-86)
struct VelGradRows {
  int nSelf, nInteriorFaces; const int* row; const int* col; const int* entryFace; const int* faceGroupOf;
  const int* groupKind; const double4* faceGeom; const double* V; const double* w; long long nnz; double* vGrad;
  FVM_DEV void operator()(long long ii) const {
    const int i = (int)ii;
    int c = i;
    bool reflect = false;
    double4 fg = make_double4(0, 0, 0, 1);
    if (i >= nSelf) {
      if (row[i + 1] - row[i] != 1) { for (int q = 0; q < 9; q++) vGrad[9 * (size_t)i + q] = 0.0; return; }
      c = col[row[i]];
      const int f = entryFace[row[i]] >> 1;
      if (f >= nInteriorFaces && groupKind[faceGroupOf[f - nInteriorFaces]] == FVMGPU_GROUP_SYMMETRY) {
        reflect = true;
        fg = faceGeom[f];
      }
    }
    double g[9];
    velGradOf(c, row, col, V, w, nnz, g);
    if (reflect) {
      const double en[3] = {fg.x / fg.w, fg.y / fg.w, fg.z / fg.w};
      double R[3][3], GT0[3][3], T1[3][3], GTR[3][3];
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          R[a][b] = (a == b) ? 1.0 - 2 * en[a] * en[b] : -2 * en[a] * en[b];
          GT0[a][b] = g[3 * b + a];  // GT0(i,j) = g0[j][i]
        }
      for (int a = 0; a < 3; a++)      // SquareTensor product: sum over k from zero, in order
        for (int b = 0; b < 3; b++) { double t = 0.0; for (int k = 0; k < 3; k++) t += R[a][k] * GT0[k][b]; T1[a][b] = t; }
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) { double t = 0.0; for (int k = 0; k < 3; k++) t += T1[a][k] * R[k][b]; GTR[a][b] = t; }
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) g[3 * b + a] = GTR[a][b];  // gr[j][i] = GTR(i,j)
    }
#pragma unroll
    for (int q = 0; q < 9; q++) vGrad[9 * (size_t)i + q] = g[q];
  }
};

// MomentumPressureGradientDiscretization, F/MomentumPressureGradientDiscretization.h:83-135:
// Green-Gauss gradient from the face pressures; ghost cells copy their neighbour's
FVM_DEV V3 pressGradOf(int c, const int* row, const int* entryFace, const double4* faceGeom, const double* pFace,
                       double vol) {
  V3 g = {0.0, 0.0, 0.0};
  for (int k = row[c]; k < row[c + 1]; k++) {
    const int ef = entryFace[k];
    const double4 fg = faceGeom[ef >> 1];
    const double pf = (ef & 1) ? -pFace[ef >> 1] : pFace[ef >> 1];
    g.x += fg.x * pf; g.y += fg.y * pf; g.z += fg.z * pf;
  }
  g.x /= vol; g.y /= vol; g.z /= vol;
  return g;
}
struct PressGradRows {
  int nSelf, nInteriorFaces; const int* row; const int* col; const int* entryFace; const int* faceGroupOf;
  const int* groupKind; const double4* faceGeom; const double4* cellGeom; const double* pFace; double* pGrad;
  FVM_DEV void operator()(long long ii) const {
    const int i = (int)ii;
    int c = i;
    bool reflect = false;
    double4 fg = make_double4(0, 0, 0, 1);
    if (i >= nSelf) {
      if (row[i + 1] - row[i] != 1) { st3(pGrad, i, V3{0, 0, 0}); return; }
      c = col[row[i]];
      const int f = entryFace[row[i]] >> 1;
      if (f >= nInteriorFaces && groupKind[faceGroupOf[f - nInteriorFaces]] == FVMGPU_GROUP_SYMMETRY) {
        reflect = true;
        fg = faceGeom[f];
      }
    }
    V3 g = pressGradOf(c, row, entryFace, faceGeom, pFace, cellGeom[c].w);
    if (reflect) {  // reflectGradient (scalar), F/GradientModel.h:21-28
      const V3 en = {fg.x / fg.w, fg.y / fg.w, fg.z / fg.w};
      const double t = 2.0 * dot3(g, en);
      g = V3{g.x - t * en.x, g.y - t * en.y, g.z - t * en.z};
    }
    st3(pGrad, i, g);
  }
};

// ---------------------------------------------------------------- momentum assembly
struct MomParams {
  int nSelf, nInteriorFaces;
  const int* row; const int* col; const int* entryFace; const int* faceGroupOf;
  const double4* cellGeom; const double4* faceGeom;
  const double* V; const double* vGrad; const double* mu; const double* rho; const double* massFlux;
  const double* contResid; const double* pGrad; const double* VN1; const double* VN2;
  const FlowBcEntry* bcs;
  double* Vnew;   // snapshot of V after the BCs (Dirichlet ghosts hold the wall velocity) = _previousVelocity
  double* diag; double* off; double* b; int* isBoundary;
  double urf; int timeOrder; double dt;
  const double* p; const double* bFaceCen;   // slip walls: cell pressure, boundary-face centroids
  double opPressure, opTemperature, molWt; int incompressible;
};

// slipJumpMomentumBC, F/FlowModelSlipJump.h:47-85: the Dirichlet value of one slip-wall face
FVM_DEV V3 slipWallVelocity(const MomParams& P, const FlowBcEntry* bc, int f, int c0) {
  const double4 fg = P.faceGeom[f];
  const V3 en = {fg.x / fg.w, fg.y / fg.w, fg.z / fg.w};
  const V3 v0 = ld3(P.V, c0);
  const double Vn = dot3(v0, en);
  const V3 Vp = {v0.x - Vn * en.x, v0.y - Vn * en.y, v0.z - Vn * en.z};
  const double4 g0 = P.cellGeom[c0];
  const double* fc = P.bFaceCen + 3 * (size_t)(f - P.nInteriorFaces);
  const V3 ds = {fc[0] - g0.x, fc[1] - g0.y, fc[2] - g0.z};
  const double dn = dot3(ds, en);
  const double pAbs = P.incompressible ? P.opPressure : (P.p[c0] + P.opPressure);
  const double Rgas = 8314.472 / P.molWt;
  const double meanFreePath = P.mu[c0] / pAbs * sqrt(0.5 * M_PI * Rgas * P.opTemperature);
  const double acc = bc->p[4];
  const double coeff = acc * meanFreePath / (dn + (acc * meanFreePath));
  const V3 Vwp = {Vp.x * coeff, Vp.y * coeff, Vp.z * coeff};
  const V3 bv = {bc->p[0], bc->p[1], bc->p[2]};
  const double bn = dot3(bv, en);
  const V3 Vwn = {bv.x - bn * bv.x, bv.y - bn * bv.y, bv.z - bn * bv.z};   // (sic) bv - dot(bv,en)*bv, :80
  return V3{Vwn.x + Vwp.x, Vwn.y + Vwp.y, Vwn.z + Vwp.z};
}

// DiffusionDiscretization<Vec3,DiagTensor3,T> for one face seen from (c0,c1)  F/DiffusionDiscretization.h:165-209
FVM_DEV void momDiffusionFace(const MomParams& P, const double4 fg, int c0, int c1, double& diffCoeff, V3& dFlux) {
  const double4 g0 = P.cellGeom[c0], g1 = P.cellGeom[c1];
  const double vol0 = g0.w, vol1 = g1.w;
  const double ds0 = g1.x - g0.x, ds1 = g1.y - g0.y, ds2 = g1.z - g0.z;
  double fd;
  if (vol0 == 0.) fd = P.mu[c1];
  else if (vol1 == 0.) fd = P.mu[c0];
  else fd = harmonicAvg(P.mu[c0], P.mu[c1]);
  const double diffMetric = fg.w * fg.w / (fg.x * ds0 + fg.y * ds1 + fg.z * ds2);
  diffCoeff = fd * diffMetric;
  const double sc0 = fd * (fg.x - ds0 * diffMetric);
  const double sc1 = fd * (fg.y - ds1 * diffMetric);
  const double sc2 = fd * (fg.z - ds2 * diffMetric);
  const double vs = vol0 + vol1;
  const double* G0 = P.vGrad + 9 * (size_t)c0;
  const double* G1 = P.vGrad + 9 * (size_t)c1;
  const V3 x0 = ld3(P.V, c0), x1 = ld3(P.V, c1);
  double sec[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {  // (gradF * secondaryCoeff)[k] = sum_i gradF[i][k] * sc_i
    const double gf0 = (G0[0 + k] * vol0 + G1[0 + k] * vol1) / vs;
    const double gf1 = (G0[3 + k] * vol0 + G1[3 + k] * vol1) / vs;
    const double gf2 = (G0[6 + k] * vol0 + G1[6 + k] * vol1) / vs;
    double s = 0.0;
    s += gf0 * sc0; s += gf1 * sc1; s += gf2 * sc2;
    sec[k] = s;
  }
  dFlux.x = diffCoeff * (x1.x - x0.x) + sec[0];
  dFlux.y = diffCoeff * (x1.y - x0.y) + sec[1];
  dFlux.z = diffCoeff * (x1.z - x0.z) + sec[2];
}

struct MomentumRows {
  MomParams P;
  FVM_DEV void operator()(long long ii) const {
    const int i = (int)ii;
    const int r0 = P.row[i], r1 = P.row[i + 1];
    const V3 xi = ld3(P.V, i);
    if (i >= P.nSelf) {
      // ghost row of a boundary face. NoSlipWall = applyDirichletBC (F/GenericBCS.h:77-115):
      // x[c1] = wall velocity, identity correction equation.
      V3 xnew = xi;
      int marks = 0;
      if (r1 - r0 == 1) {
        const int f = P.entryFace[r0] >> 1;
        if (f >= P.nInteriorFaces) {
          const FlowBcEntry* bc = faceBc(P.bcs, P.faceGroupOf, P.nInteriorFaces, f);
          const bool inflowOutflow = bc->kind == FVMGPU_FLOWBC_VELOCITY || bc->kind == FVMGPU_FLOWBC_PRESSURE;
          if (inflowOutflow && P.massFlux[f] > 0.) {
            // applyExtrapolationBC (F/GenericBCS.h:180-212): identity-like boundary row x1 = x0
            const V3 x0 = ld3(P.V, P.col[r0]);
            st3(P.Vnew, i, xi);
            st3(P.diag, i, V3{-1.0, -1.0, -1.0});
            st3(P.b, i, V3{x0.x - xi.x, x0.y - xi.y, x0.z - xi.z});
            P.off[r0] = 1.0;
            P.isBoundary[i] = 1;
            return;
          }
          if (bc->kind == FVMGPU_FLOWBC_NOSLIP_WALL || inflowOutflow) xnew = V3{bc->p[0], bc->p[1], bc->p[2]};
          else if (bc->kind == FVMGPU_FLOWBC_SLIP_JUMP) xnew = slipWallVelocity(P, bc, f, P.col[r0]);
          else if (bc->kind == FVMGPU_FLOWBC_SYMMETRY) {  // x[c1] = x[c0] - 2 (x[c0].en) en, F/GenericBCS.h:583-596
            const double4 fg = P.faceGeom[f];
            const V3 en = {fg.x / fg.w, fg.y / fg.w, fg.z / fg.w};
            const V3 x0 = ld3(P.V, P.col[r0]);
            const double d = dot3(x0, en);
            xnew = V3{x0.x - 2 * d * en.x, x0.y - 2 * d * en.y, x0.z - 2 * d * en.z};
            marks = 1;  // setBoundary(c1); its diagonal (= the neighbour's, :603) is written by the neighbour's thread
          }
        }
        P.off[r0] = 0.0;
      }
      st3(P.Vnew, i, xnew);
      if (!marks) st3(P.diag, i, V3{-1.0, -1.0, -1.0});
      st3(P.b, i, V3{0.0, 0.0, 0.0});
      P.isBoundary[i] = marks;
      return;
    }
    double diag = 0.0;          // common part of the DiagonalTensor diagonal
    V3 dpd = {0.0, 0.0, 0.0};   // per-component part (fixedPressureMomentumBC)
    V3 r = {0.0, 0.0, 0.0};
    bool hasB = false;
    // ---- DiffusionDiscretization (all faces in face order)
    for (int k = r0; k < r1; k++) {
      const int ef = P.entryFace[k];
      const int f = ef >> 1, side = ef & 1;
      if (f >= P.nInteriorFaces) hasB = true;
      const double4 fg = P.faceGeom[f];
      double dc; V3 df;
      if (side == 0) { momDiffusionFace(P, fg, i, P.col[k], dc, df); r.x += df.x; r.y += df.y; r.z += df.z; }
      else { momDiffusionFace(P, fg, P.col[k], i, dc, df); r.x -= df.x; r.y -= df.y; r.z -= df.z; }
      P.off[k] = dc;
      diag -= dc;
    }
    // ---- ConvectionDiscretization (upwind, F/ConvectionDiscretization.h:166-199)
    for (int k = r0; k < r1; k++) {
      const int ef = P.entryFace[k];
      const int f = ef >> 1, side = ef & 1;
      const double flux = P.massFlux[f];
      const V3 xo = ld3(P.V, P.col[k]);
      const V3 x0 = side ? xo : xi, x1 = side ? xi : xo;
      const V3 up = (flux > 0.0) ? x0 : x1;
      const V3 vf = {flux * up.x, flux * up.y, flux * up.z};
      if (side == 0) {
        if (flux > 0.0) diag -= flux; else P.off[k] -= flux;
        r.x -= vf.x; r.y -= vf.y; r.z -= vf.z;
      } else {
        if (flux > 0.0) P.off[k] += flux; else diag += flux;
        r.x += vf.x; r.y += vf.y; r.z += vf.z;
      }
    }
    diag += P.contResid[i];
    // ---- MomentumPressureGradientDiscretization :124-130
    const double vol = P.cellGeom[i].w;
    {
      const V3 pg = ld3(P.pGrad, i);
      r.x -= vol * pg.x; r.y -= vol * pg.y; r.z -= vol * pg.z;
    }
    // ---- TimeDerivativeDiscretization (static mesh), F/TimeDerivativeDiscretization.h:102-108,149-155
    if (P.timeOrder == 1) {
      const double rhoVbydT = P.rho[i] * vol / P.dt;
      const V3 n1 = ld3(P.VN1, i);
      r.x -= rhoVbydT * (xi.x - n1.x); r.y -= rhoVbydT * (xi.y - n1.y); r.z -= rhoVbydT * (xi.z - n1.z);
      diag -= rhoVbydT;
    } else if (P.timeOrder == 2) {
      const double rhoVbydT = P.rho[i] * vol / P.dt;
      const V3 n1 = ld3(P.VN1, i), n2 = ld3(P.VN2, i);
      r.x -= rhoVbydT * (1.5 * xi.x - 2.0 * n1.x + 0.5 * n2.x);
      r.y -= rhoVbydT * (1.5 * xi.y - 2.0 * n1.y + 0.5 * n2.y);
      r.z -= rhoVbydT * (1.5 * xi.z - 2.0 * n1.z + 0.5 * n2.z);
      diag -= rhoVbydT * 1.5;
    }
    // ---- BC loop (NoSlipWall -> applyDirichletBC): r[c0] += coeff01 * (bValue - x[c1]); coeff01 = 0
    if (hasB) {
      for (int k = r0; k < r1; k++) {
        const int f = P.entryFace[k] >> 1;
        if (f < P.nInteriorFaces) continue;
        const FlowBcEntry* bc = faceBc(P.bcs, P.faceGroupOf, P.nInteriorFaces, f);
        const V3 x1 = ld3(P.V, P.col[k]);
        const double c01 = P.off[k];
        const bool inflowOutflow = bc->kind == FVMGPU_FLOWBC_VELOCITY || bc->kind == FVMGPU_FLOWBC_PRESSURE;
        if (inflowOutflow && P.massFlux[f] > 0.) {
          // applyExtrapolationBC: dFluxdXC1 = -diag[c1]; for an outflow face the ghost row's diagonal is
          // -diffCoeff (upwind convection adds nothing to it) and coeff01 is +diffCoeff
          const double dFluxdXC1 = c01;
          diag += dFluxdXC1;
          r.x += dFluxdXC1 * (xi.x - x1.x); r.y += dFluxdXC1 * (xi.y - x1.y); r.z += dFluxdXC1 * (xi.z - x1.z);
          P.off[k] = 0.0;
        } else if (bc->kind == FVMGPU_FLOWBC_NOSLIP_WALL || inflowOutflow) {
          r.x += c01 * (bc->p[0] - x1.x); r.y += c01 * (bc->p[1] - x1.y); r.z += c01 * (bc->p[2] - x1.z);
          P.off[k] = 0.0;
        } else if (bc->kind == FVMGPU_FLOWBC_SLIP_JUMP) {  // applyDirichletBC(f, Vw)
          const V3 vw = slipWallVelocity(P, bc, f, i);
          r.x += c01 * (vw.x - x1.x); r.y += c01 * (vw.y - x1.y); r.z += c01 * (vw.z - x1.z);
          P.off[k] = 0.0;
        } else if (bc->kind == FVMGPU_FLOWBC_SYMMETRY) {  // applySymmetryBC, F/GenericBCS.h:569-615
          const double4 fg = P.faceGeom[f];
          const V3 en = {fg.x / fg.w, fg.y / fg.w, fg.z / fg.w};
          const double d = dot3(xi, en);
          const V3 xB = {xi.x - 2 * d * en.x, xi.y - 2 * d * en.y, xi.z - 2 * d * en.z};
          r.x += c01 * (xB.x - x1.x); r.y += c01 * (xB.y - x1.y); r.z += c01 * (xB.z - x1.z);
          P.off[k] = 0.0;
          st3(P.diag, P.col[k], V3{diag, diag, diag});  // _dRdXDiag[c1] = _dRdXDiag[c0] (before the under-relaxation)
        }
      }
      // fixedPressureMomentumBC (F/FlowModelPressureBC.h:11-50): inflow faces of pressure boundaries
      for (int k = r0; k < r1; k++) {
        const int f = P.entryFace[k] >> 1;
        if (f < P.nInteriorFaces) continue;
        const FlowBcEntry* bc = faceBc(P.bcs, P.faceGroupOf, P.nInteriorFaces, f);
        if (bc->kind != FVMGPU_FLOWBC_PRESSURE || !(P.massFlux[f] < 0.)) continue;
        const double4 fg = P.faceGeom[f];
        const V3 vb = {bc->p[0], bc->p[1], bc->p[2]};   // V[c1] after the Dirichlet BC of this face
        const double dpdV = -P.rho[i] * dot3(vb, vb) / P.urf;
        dpd.x += dpdV * fg.x * fg.x / fg.w; dpd.y += dpdV * fg.y * fg.y / fg.w; dpd.z += dpdV * fg.z * fg.z / fg.w;
      }
    }
    // ---- Underrelaxer, F/Underrelaxer.h:49-52
    const V3 dg = {(diag + dpd.x) / P.urf, (diag + dpd.y) / P.urf, (diag + dpd.z) / P.urf};
    st3(P.Vnew, i, xi);
    st3(P.diag, i, dg);
    st3(P.b, i, r);
    P.isBoundary[i] = 0;
  }
};

struct SplitComponent {  // scalar system of velocity component k
  int k; const double* diag3; const double* b3; double* diag; double* b; double* delta;
  FVM_DEV void operator()(long long i) const { diag[i] = diag3[3 * i + k]; b[i] = b3[3 * i + k]; delta[i] = 0.0; }
};
struct DiffCountRows {  // number of rows whose diagonal differs from the one the hierarchy was built for
  const double* a; const double* b;
  FVM_DEV void operator()(long long i, double* o) const { o[0] = (a[i] != b[i]) ? 1.0 : 0.0; }
};
struct AbsRows3 {  // 1-norms of the three components of an AoS vector field
  const double* a;
  FVM_DEV void operator()(long long i, double* o) const { o[0] = fabs(a[3 * i]); o[1] = fabs(a[3 * i + 1]); o[2] = fabs(a[3 * i + 2]); }
};
struct MergeComponent {  // delta3[.][k] = delta ; V[.][k] += delta  (ls.updateSolution)
  int k; const double* delta; double* delta3; double* V;
  FVM_DEV void operator()(long long i) const { const double d = delta[i]; delta3[3 * i + k] = d; V[3 * i + k] += d; }
};
struct CopyRows { const double* a; double* b; FVM_DEV void operator()(long long i) const { b[i] = a[i]; } };
// CRMatrix::solveBoundary for the marked ghost rows of the momentum system (extrapolation boundaries),
// F/CRMatrix.h:433-454, then x += delta
struct MomentumGhostRows {
  int nSelf; const int* row; const int* col; const int* isBoundary; const double* diag3; const double* off;
  const double* b3; double* delta3; double* V;
  FVM_DEV void operator()(long long k) const {
    const int i = nSelf + (int)k;
    if (!isBoundary[i]) return;
    const int q = row[i];
    if (row[i + 1] - q != 1) return;
    const int c0 = col[q];
    for (int c = 0; c < 3; c++) {
      const double d = -(b3[3 * (size_t)i + c] + off[q] * delta3[3 * (size_t)c0 + c]) / diag3[3 * (size_t)i + c];
      delta3[3 * (size_t)i + c] = d;
      V[3 * (size_t)i + c] += d;
    }
  }
};

// ---------------------------------------------------------------- continuity
struct ContParams {
  int nSelf, nInteriorFaces;
  const int* faceCells; const int* row; const int* col; const int* entryFace; const int* faceGroupOf;
  const double4* cellGeom; const double4* faceGeom;
  const double* V; const double* Vprev; const double* momAp; const double* p; const double* pGrad; const double* rho;
  const FlowBcEntry* bcs;
  double* massFlux; double* pCoeff;
  double* bDiagAdd; double* bOff10; double* bCoeffL;   // per boundary face (index f - nInteriorFaces)
  double urf;
};

// discretizeMassFluxInterior for one interior face, F/FlowModelInterior.h:68-117
struct MassFluxFaces {
  ContParams P;
  FVM_DEV void operator()(long long ff) const {
    const int f = (int)ff;
    const int c0 = P.faceCells[2 * f], c1 = P.faceCells[2 * f + 1];
    const double4 fg = P.faceGeom[f];
    const V3 Af = {fg.x, fg.y, fg.z};
    if (!interiorLike(P.bcs, P.faceGroupOf, P.nInteriorFaces, f)) {
      const FlowBcEntry* bc = faceBc(P.bcs, P.faceGroupOf, P.nInteriorFaces, f);
      const int bf = f - P.nInteriorFaces;
      P.pCoeff[f] = 0.0;
      if (bc->kind != FVMGPU_FLOWBC_PRESSURE) {  // fixedFluxContinuityBC, F/FlowModelVelocityBC.h:64-92
        const V3 bv = {bc->p[0], bc->p[1], bc->p[2]};
        P.massFlux[f] = P.rho[c0] * dot3(bv, Af);
        P.bDiagAdd[bf] = 0.0; P.bOff10[bf] = 1.0; P.bCoeffL[bf] = 0.0;
        return;
      }
      // fixedPressureContinuityBC, F/FlowModelPressureBC.h:104-158
      const double4 g0 = P.cellGeom[c0], g1 = P.cellGeom[c1];
      const V3 ds = {g1.x - g0.x, g1.y - g0.y, g1.z - g0.z};
      const double dpf = dot3(ld3(P.pGrad, c0), ds) - P.p[c1] + P.p[c0];
      const double rhoF = P.rho[c0];
      const V3 a0 = ld3(P.momAp, c0);
      const double Q = rhoF * (Af.x * Af.x / a0.x + Af.y * Af.y / a0.y + Af.z * Af.z / a0.z) * g0.w / dot3(Af, ds);
      const double oneMinusUrf = 1.0 - P.urf;
      const double massFluxI = rhoF * (dot3(ld3(P.V, c0), Af) - oneMinusUrf * dot3(ld3(P.Vprev, c0), Af)) - Q * dpf +
                               oneMinusUrf * P.massFlux[f];
      const V3 Vb = ld3(P.V, c1);
      const double massFluxB = rhoF * dot3(Vb, Af);
      P.massFlux[f] = massFluxI;
      double Vb_dpdVb = 0.0;
      if (massFluxB < 0) Vb_dpdVb = -dot3(Vb, Vb) * rhoF;
      const double denom = massFluxI - Q * Vb_dpdVb;
      if (denom != 0) {
        const double dMassFluxdp0 = -Q * massFluxI / denom;
        P.bCoeffL[bf] = dMassFluxdp0;
        P.bDiagAdd[bf] = -dMassFluxdp0;          // ppDiag[c0] -= dMassFluxdp0
        P.bOff10[bf] = -Q * Vb_dpdVb / denom;    // coeff10 = dpbdp0
      } else {                                   // treat as fixed pressure
        P.bCoeffL[bf] = -Q;
        P.bDiagAdd[bf] = Q;
        P.bOff10[bf] = 0.0;
      }
      return;
    }
    const double4 g0 = P.cellGeom[c0], g1 = P.cellGeom[c1];
    const V3 ds = {g1.x - g0.x, g1.y - g0.y, g1.z - g0.z};
    const double Ads = dot3(Af, ds);
    const double diffMetric = fg.w * fg.w / Ads;
    const V3 a0 = ld3(P.momAp, c0), a1 = ld3(P.momAp, c1);
    const double momApBar0 = (a0.x + a0.y + a0.z) / 3.0;
    const double momApBar1 = (a1.x + a1.y + a1.z) / 3.0;
    const double momApBarFace = momApBar0 + momApBar1;
    const double oneMinusUrf = 1.0 - P.urf;
    const double VdotA0 = dot3(ld3(P.V, c0), Af) - oneMinusUrf * dot3(ld3(P.Vprev, c0), Af);
    const double VdotA1 = dot3(ld3(P.V, c1), Af) - oneMinusUrf * dot3(ld3(P.Vprev, c1), Af);
    const double dpf = g0.w * dot3(ld3(P.pGrad, c0), ds) + g1.w * dot3(ld3(P.pGrad, c1), ds);
    const double Vn = (VdotA0 * momApBar0 + VdotA1 * momApBar1 - dpf * diffMetric) / momApBarFace;
    const double rhoF = 0.5 * (P.rho[c0] + P.rho[c1]);
    const double aByMomAp = Af.x * Af.x / (a0.x + a1.x) + Af.y * Af.y / (a0.y + a1.y) + Af.z * Af.z / (a0.z + a1.z);
    const double pCoeff = rhoF * aByMomAp * (g0.w + g1.w) / Ads;
    P.massFlux[f] = rhoF * Vn - pCoeff * (P.p[c0] - P.p[c1]) + (1 - P.urf) * P.massFlux[f];
    P.pCoeff[f] = pCoeff;
  }
};

struct BoundaryFluxSum {  // netFlux over the boundary faces (not the partition interfaces); volume of the interior cells
  int nInteriorFaces; const double* massFlux; const FlowBcEntry* bcs; const int* faceGroupOf;
  FVM_DEV void operator()(long long k, double* o) const {
    o[0] = bcs[faceGroupOf[k]].groupKind == FVMGPU_GROUP_INTERFACE ? 0.0 : massFlux[nInteriorFaces + k];
  }
};
struct VolumeSum { const double4* cellGeom; FVM_DEV void operator()(long long i, double* o) const { o[0] = cellGeom[i].w; } };

struct ContinuityRows {  // pressure-correction matrix rows, F/FlowModelInterior.h:105-117 + BCs + :1144-1188
  ContParams P;
  const double* scal;  // [0] netFlux, [1] volumeSum
  int useReferencePressure, refCell;
  double* diag; double* off; double* b; int* isBoundary;
  FVM_DEV void operator()(long long ii) const {
    const int i = (int)ii;
    const int r0 = P.row[i], r1 = P.row[i + 1];
    if (i >= P.nSelf) {
      // fixedFlux / fixedPressure ContinuityBC on the ghost: ppDiag = -1, r = 0, coeff10 = 1 (fixed flux) or
      // dpbdp0 (pressure boundary), setBoundary
      for (int k = r0; k < r1; k++) {
        const int f = P.entryFace[k] >> 1;
        off[k] = f >= P.nInteriorFaces ? P.bOff10[f - P.nInteriorFaces] : 1.0;
      }
      diag[i] = -1.0;
      b[i] = 0.0;
      isBoundary[i] = 1;
      return;
    }
    double d = 0.0, r = 0.0;
    for (int k = r0; k < r1; k++) {
      const int ef = P.entryFace[k];
      const int f = ef >> 1;
      const double mf = P.massFlux[f];
      if (interiorLike(P.bcs, P.faceGroupOf, P.nInteriorFaces, f)) {
        const double pc = P.pCoeff[f];
        if (ef & 1) r += mf; else r -= mf;
        off[k] = -pc;
        d += pc;
      } else {
        r -= mf;       // this row is c0 of the boundary face
        off[k] = 0.0;  // coeff01 = 0
        d += P.bDiagAdd[f - P.nInteriorFaces];
      }
    }
    if (useReferencePressure) {
      r += (scal[0] / scal[1]) * P.cellGeom[i].w;
      if (i == refCell) {  // setDirichlet, F/FlowModel_impl.h:970-975
        d = -1.0;
        r = 0.0;
        for (int k = r0; k < r1; k++) off[k] = 0.0;
      }
    }
    diag[i] = d;
    b[i] = r;
    isBoundary[i] = 0;
  }
};

// ---- after the pressure-correction solve
struct PpGhostRows {  // CRMatrix::solveBoundary for the marked ghost rows, F/CRMatrix.h:433-454
  int nSelf; const int* row; const int* col; const double* diag; const double* off; const double* b; double* pp;
  FVM_DEV void operator()(long long k) const {
    const int i = nSelf + (int)k;
    double sum = b[i];
    for (int q = row[i]; q < row[i + 1]; q++) sum += off[q] * pp[col[q]];
    pp[i] = -sum / diag[i];
  }
};
struct CorrectPressureRows {  // correctPressure, F/FlowModel_impl.h:844-861
  const double* pp; const double* refPP; double urf; double* p;
  FVM_DEV void operator()(long long i) const { p[i] += urf * (pp[i] - (refPP ? refPP[0] : 0.0)); }
};
struct CorrectMassFluxFaces {  // correctMassFluxInterior, F/FlowModelInterior.h:390-399
  const int* faceCells; const int* pairToCol; const double* off; const double* pp; double* massFlux;
  FVM_DEV void operator()(long long ff) const {
    const int f = (int)ff;
    const int c0 = faceCells[2 * f], c1 = faceCells[2 * f + 1];
    massFlux[f] -= off[pairToCol[2 * f]] * pp[c1] - off[pairToCol[2 * f + 1]] * pp[c0];
  }
};
// the same for the partition-interface faces: coeff01 = coeff10 = -pCoeff (the ghost row of the pressure-
// correction matrix is not assembled here, F/FlowModelInterior.h:105-117)
struct CorrectInterfaceMassFluxFaces {
  int nInteriorFaces; const int* faceCells; const FlowBcEntry* bcs; const int* faceGroupOf; const double* pCoeff;
  const double* pp; double* massFlux;
  FVM_DEV void operator()(long long k) const {
    if (bcs[faceGroupOf[k]].groupKind != FVMGPU_GROUP_INTERFACE) return;
    const int f = nInteriorFaces + (int)k;
    const int c0 = faceCells[2 * f], c1 = faceCells[2 * f + 1];
    const double c = -pCoeff[f];
    massFlux[f] -= c * pp[c1] - c * pp[c0];
  }
};
struct ReferencePpKernel { const double* pp; int refCell; double* out; FVM_DEV void operator()(long long) const { out[0] = refCell >= 0 ? pp[refCell] : 0.0; } };
// coefficients of the face pressure interpolation, F/FlowModelInterior.h:252-266, 335-349
FVM_DEV void facePressureWeights(const ContParams& P, int f, int c0, int c1, double& coeff0, double& coeff1) {
  const double4 fg = P.faceGeom[f];
  const double4 g0 = P.cellGeom[c0], g1 = P.cellGeom[c1];
  const V3 ds = {g1.x - g0.x, g1.y - g0.y, g1.z - g0.z};
  const V3 Af = {fg.x, fg.y, fg.z};
  const V3 a0 = ld3(P.momAp, c0), a1 = ld3(P.momAp, c1);
  const double aBy0 = Af.x * Af.x / a0.x + Af.y * Af.y / a0.y + Af.z * Af.z / a0.z;
  const double aBy1 = Af.x * Af.x / a1.x + Af.y * Af.y / a1.y + Af.z * Af.z / a1.z;
  const double Adotes = dot3(Af, ds) / sqrt(dot3(ds, ds));
  coeff0 = g0.w * P.rho[c0] * aBy0 / Adotes;
  coeff1 = g1.w * P.rho[c1] * aBy1 / Adotes;
}
struct CorrectVelocityRows {  // correctVelocityInterior + correctVelocityBoundary as a per-cell gather
  ContParams P; const double* pp; double* Vout;
  FVM_DEV void operator()(long long ii) const {
    const int i = (int)ii;
    V3 v = ld3(P.V, i);
    const V3 ap = ld3(P.momAp, i);
    for (int k = P.row[i]; k < P.row[i + 1]; k++) {
      const int ef = P.entryFace[k];
      const int f = ef >> 1, side = ef & 1;
      const double4 fg = P.faceGeom[f];
      if (interiorLike(P.bcs, P.faceGroupOf, P.nInteriorFaces, f)) {
        const int c0 = side ? P.col[k] : i, c1 = side ? i : P.col[k];
        double w0, w1;
        facePressureWeights(P, f, c0, c1, w0, w1);
        const double ppFace = (w0 * pp[c0] + w1 * pp[c1]) / (w0 + w1);
        const V3 ppA = {ppFace * fg.x, ppFace * fg.y, ppFace * fg.z};
        if (side == 0) { v.x += ppA.x / ap.x; v.y += ppA.y / ap.y; v.z += ppA.z / ap.z; }
        else { v.x -= ppA.x / ap.x; v.y -= ppA.y / ap.y; v.z -= ppA.z / ap.z; }
      } else {  // correctVelocityBoundary, F/FlowModel_impl.h:803-828
        const double ppFace = pp[P.col[k]];
        v.x += ppFace * fg.x / ap.x; v.y += ppFace * fg.y / ap.y; v.z += ppFace * fg.z / ap.z;
      }
    }
    st3(Vout, i, v);
  }
};
struct CorrectBoundaryMassFluxFaces {  // flux row of a boundary face: dMassFlux = coeffL * pp[c0] (coeffR = 0, rFlux = 0)
  int nInteriorFaces; const int* faceCells; const double* coeffL; const double* pp; double* massFlux;
  FVM_DEV void operator()(long long k) const {
    const int f = nInteriorFaces + (int)k;
    massFlux[f] -= coeffL[k] * pp[faceCells[2 * f]];
  }
};
// pressureBoundaryPostContinuitySolve, F/FlowModelPressureBC.h:161-212.
// NB (reference ordering, reproduced): the reference treats the boundary groups one after the other
// (correctVelocityBoundary, then this function, F/FlowModel_impl.h:1309-1330), so the neighbour's
// velocity copied into an outflow ghost does not yet contain the velocity corrections of the boundary
// faces in LATER groups (a corner cell touching the outlet and a wall). Here all corrections are
// already applied, so those later contributions are taken out again.
struct PressureBoundaryPostFaces {
  ContParams P; const double* pp; double* V; double* p;
  FVM_DEV void operator()(long long k) const {
    const int f = P.nInteriorFaces + (int)k;
    const FlowBcEntry* bc = faceBc(P.bcs, P.faceGroupOf, P.nInteriorFaces, f);
    if (bc->kind != FVMGPU_FLOWBC_PRESSURE) return;
    const int c0 = P.faceCells[2 * f], c1 = P.faceCells[2 * f + 1];
    const double bp = bc->p[3];
    const double rhoF = P.rho[c0];
    const double4 fg = P.faceGeom[f];
    if (P.massFlux[f] > 0) {
      V3 v = ld3(V, c0);
      const int myGroup = P.faceGroupOf[f - P.nInteriorFaces];
      const V3 ap = ld3(P.momAp, c0);
      for (int q = P.row[c0]; q < P.row[c0 + 1]; q++) {
        const int f2 = P.entryFace[q] >> 1;
        if (interiorLike(P.bcs, P.faceGroupOf, P.nInteriorFaces, f2) || P.faceGroupOf[f2 - P.nInteriorFaces] <= myGroup)
          continue;  // (interfaces are corrected with the interior faces, before every boundary group)
        const double4 g2 = P.faceGeom[f2];
        const double ppFace = pp[P.col[q]];
        v.x -= ppFace * g2.x / ap.x; v.y -= ppFace * g2.y / ap.y; v.z -= ppFace * g2.z / ap.z;
      }
      st3(V, c1, v);
      p[c1] = bp;
    } else {
      const double Vn = -P.massFlux[f] / (rhoF * fg.w);
      const V3 v = {-Vn * fg.x / fg.w, -Vn * fg.y / fg.w, -Vn * fg.z / fg.w};
      st3(V, c1, v);
      p[c1] = bp - 0.5 * rhoF * dot3(v, v);
    }
  }
};
struct FacePressureFaces {  // updateFacePressureInterior / Boundary
  ContParams P; double* pFace;
  FVM_DEV void operator()(long long ff) const {
    const int f = (int)ff;
    const int c0 = P.faceCells[2 * f], c1 = P.faceCells[2 * f + 1];
    if (!interiorLike(P.bcs, P.faceGroupOf, P.nInteriorFaces, f)) { pFace[f] = P.p[c1]; return; }
    double w0, w1;
    facePressureWeights(P, f, c0, c1, w0, w1);
    pFace[f] = (w0 * P.p[c0] + w1 * P.p[c1]) / (w0 + w1);
  }
};
struct FillRows { double* p; double v; FVM_DEV void operator()(long long i) const { p[i] = v; } };

// ================================================================= host side
static System* makeScalarSystem(Mesh* m, bool multi) {
  std::unique_ptr<System> s(new System);
  // solved as a stand-alone CSR system on the mesh's cellCells pattern; a mesh part keeps its mesh so that
  // the solver finds the halo maps (interface ghost columns, distributed hierarchy)
  s->mesh = multi ? m : nullptr;
  s->nSelf = m->nSelf; s->nTotal = m->nTotal; s->nnz = m->nnz;
  s->row = m->row.p; s->col = m->col.p;
  const size_t nt = (size_t)m->nTotal;
  s->diag.alloc(nt); s->b.alloc(nt); s->delta.alloc(nt); s->x.alloc(nt); s->off.alloc((size_t)m->nnz);
  s->isBoundary.alloc(nt);
  s->diag.zero(); s->b.zero(); s->delta.zero(); s->x.zero(); s->off.zero(); s->isBoundary.zero();
  s->version = nextVersion();
  s->patternVersion = nextVersion();
  return s.release();
}

Flow* flowCreate(Mesh* m) {
  requireReady();
  if (!m->hasGeometry) fail("flow: mesh geometry not set (fvmgpu_mesh_set_geometry)");
  std::unique_ptr<Flow> F(new Flow);
  F->mesh = m;
  F->multi = commActive();  // every rank takes part in the all-reduces; an empty halo exchanges nothing
  F->nSelf = m->nSelf; F->nTotal = m->nTotal; F->nFaces = m->nFaces; F->nnz = m->nnz;
  const size_t nt = (size_t)m->nTotal, nf = (size_t)m->nFaces;
  F->V.alloc(3 * nt); F->Vprev.alloc(3 * nt); F->p.alloc(nt); F->pFace.alloc(nf); F->rho.alloc(nt); F->mu.alloc(nt);
  F->massFlux.alloc(nf); F->contResid.alloc(nt); F->pGrad.alloc(3 * nt); F->vGrad.alloc(9 * nt); F->momAp.alloc(3 * nt);
  F->mDiag.alloc(3 * nt); F->mOff.alloc((size_t)m->nnz); F->mB.alloc(3 * nt); F->mDelta.alloc(3 * nt);
  F->pCoeff.alloc(nf);
  {
    const size_t nb = nf - (size_t)m->nInteriorFaces + 1;
    F->bDiagAdd.alloc(nb); F->bOff10.alloc(nb); F->bCoeffL.alloc(nb);
    F->bDiagAdd.zero(); F->bOff10.zero(); F->bCoeffL.zero();
  }
  F->V.zero(); F->Vprev.zero(); F->p.zero(); F->pFace.zero(); F->massFlux.zero(); F->contResid.zero();
  F->pGrad.zero(); F->vGrad.zero(); F->momAp.zero(); F->mDiag.zero(); F->mOff.zero(); F->mB.zero(); F->mDelta.zero();
  F->pCoeff.zero();
  parallelFor((long long)nt, FillRows{F->rho.p, 1.0});
  parallelFor((long long)nt, FillRows{F->mu.p, 1e-3});
  F->comp.reset(makeScalarSystem(m, F->multi));
  F->pp.reset(makeScalarSystem(m, F->multi));
  F->lastDiag.alloc(nt);
  F->scal.alloc(16);
  F->scal.zero();
  for (const FaceGroup& g : m->groups) {
    FlowBcEntry e;
    e.offset = g.offset; e.count = g.count; e.kind = -1; e.groupKind = g.kind;
    e.p[0] = e.p[1] = e.p[2] = e.p[3] = 0.0; e.p[4] = 1.0;
    F->bcs.push_back(e);
  }
  streamSync();
  return F.release();
}

static DBuf<double>* flowField(Flow* F, int field, size_t& len) {
  const size_t nt = (size_t)F->nTotal, nf = (size_t)F->nFaces;
  switch (field) {
    case FVMGPU_FLOW_VELOCITY: len = 3 * nt; return &F->V;
    case FVMGPU_FLOW_PRESSURE: len = nt; return &F->p;
    case FVMGPU_FLOW_DENSITY: len = nt; return &F->rho;
    case FVMGPU_FLOW_VISCOSITY: len = nt; return &F->mu;
    case FVMGPU_FLOW_MASS_FLUX: len = nf; return &F->massFlux;
    case FVMGPU_FLOW_FACE_PRESSURE: len = nf; return &F->pFace;
    case FVMGPU_FLOW_PRESSURE_GRADIENT: len = 3 * nt; return &F->pGrad;
    case FVMGPU_FLOW_VELOCITY_GRADIENT: len = 9 * nt; return &F->vGrad;
    case FVMGPU_FLOW_CONT_RESID: len = nt; return &F->contResid;
    case FVMGPU_FLOW_MOM_AP: len = 3 * nt; return &F->momAp;
    case FVMGPU_FLOW_PREV_VELOCITY: len = 3 * nt; return &F->Vprev;
    case FVMGPU_FLOW_VELOCITY_N1: len = 3 * nt; return &F->VN1;
    case FVMGPU_FLOW_VELOCITY_N2: len = 3 * nt; return &F->VN2;
    default: return nullptr;
  }
}
void flowSetField(Flow* F, int field, const double* host, long long n, bool fill, double value) {
  requireReady();
  size_t len = 0;
  DBuf<double>* buf = flowField(F, field, len);
  if (!buf) fail("flow_set_field: unknown field %d", field);
  if (!fill && (size_t)n != len) fail("flow_set_field: field %d expects %zu values, got %lld", field, len, n);
  if (buf->n < len) buf->alloc(len);
  if (fill) parallelFor((long long)len, FillRows{buf->p, value});
  else buf->upload(host, len);
  if (field == FVMGPU_FLOW_MOM_AP) F->hasMomAp = true;
  if (field == FVMGPU_FLOW_VELOCITY_N1) F->hasVN1 = true;
  if (field == FVMGPU_FLOW_VELOCITY_N2) F->hasVN2 = true;
}
void flowGetField(Flow* F, int field, double* host, long long n) {
  requireReady();
  size_t len = 0;
  DBuf<double>* buf = flowField(F, field, len);
  if (!buf || !buf->p) fail("flow_get_field: field %d not available", field);
  if ((size_t)n != len) fail("flow_get_field: field %d has %zu values, asked for %lld", field, len, n);
  buf->download(host, len);
}
void flowSetBc(Flow* F, int groupId, int kind, const double* p, int np) {
  requireReady();
  if (kind < FVMGPU_FLOWBC_NOSLIP_WALL || kind > FVMGPU_FLOWBC_SLIP_JUMP) fail("flow_set_bc: unknown boundary kind %d", kind);
  if (kind == FVMGPU_FLOWBC_SLIP_JUMP && !F->mesh->bFaceCen.p)
    fail("flow_set_bc: SlipJump needs the face centroids (pass them to fvmgpu_mesh_set_geometry)");
  for (size_t g = 0; g < F->bcs.size(); g++) {
    const FaceGroup& fg = F->mesh->groups[g];
    if (fg.id == groupId && fg.kind != FVMGPU_GROUP_INTERIOR) {
      FlowBcEntry& e = F->bcs[g];
      e.kind = kind;
      for (int i = 0; i < 5; i++) e.p[i] = (i < np && p) ? p[i] : (i == 4 ? 1.0 : 0.0);
      F->bcsDirty = true;
      return;
    }
  }
  fail("flow_set_bc: no boundary face group with id %d", groupId);
}
static void flowSyncBcs(Flow* F) {
  if (!F->bcsDirty) return;
  for (size_t g = 1; g < F->bcs.size(); g++)
    if (F->bcs[g].kind < 0 && F->bcs[g].groupKind != FVMGPU_GROUP_INTERFACE)
      fail("flow: boundary group %d has no boundary condition", F->mesh->groups[g].id);
  F->bcsDev.upload(F->bcs.data(), F->bcs.size());
  F->hasPressureBoundary = false;
  for (const FlowBcEntry& e : F->bcs) if (e.kind == FVMGPU_FLOWBC_PRESSURE) F->hasPressureBoundary = true;
  if (!F->multi) F->pressureBoundaryAnywhere = F->hasPressureBoundary;  // across ranks: agreed in flowInit
  F->bcsDirty = false;
}
static ContParams contParams(Flow* F, double urf) {
  Mesh* m = F->mesh;
  ContParams P;
  P.nSelf = m->nSelf; P.nInteriorFaces = m->nInteriorFaces;
  P.faceCells = m->faceCells.p; P.row = m->row.p; P.col = m->col.p; P.entryFace = m->entryFace.p;
  P.faceGroupOf = m->faceGroupOf.p; P.cellGeom = m->cellGeom.p; P.faceGeom = m->faceGeom.p;
  P.V = F->V.p; P.Vprev = F->Vprev.p; P.momAp = F->momAp.p; P.p = F->p.p; P.pGrad = F->pGrad.p; P.rho = F->rho.p;
  P.bcs = F->bcsDev.p; P.massFlux = F->massFlux.p; P.pCoeff = F->pCoeff.p; P.urf = urf;
  P.bDiagAdd = F->bDiagAdd.p; P.bOff10 = F->bOff10.p; P.bCoeffL = F->bCoeffL.p;
  return P;
}
// MultiField::syncLocal for one cell field (width doubles per cell, AoS)
static void flowExchange(Flow* F, DBuf<double>& field, int width) {
  if (F->multi) F->mesh->halo.exchange(field.p, width);
}
static void flowContinuityResidual(Flow* F) {
  Mesh* m = F->mesh;
  parallelFor(m->nTotal, ContResidRows{m->row.p, m->entryFace.p, F->massFlux.p, F->contResid.p});
}

// FlowModel::init: default face mass fluxes + continuity residual (F/FlowModel_impl.h:222-340)
void flowInit(Flow* F) {
  requireReady();
  flowSyncBcs(F);
  Mesh* m = F->mesh;
  F->pressureBoundaryAnywhere = F->multi ? commAny(F->hasPressureBoundary) : F->hasPressureBoundary;
  // ghost copies of the fields the host uploaded (the reference's arrays arrive synced from the partitioner)
  flowExchange(F, F->V, 3); flowExchange(F, F->p, 1); flowExchange(F, F->rho, 1); flowExchange(F, F->mu, 1);
  parallelFor(m->nFaces, FlowInitMassFluxFaces{m->nInteriorFaces, m->faceCells.p, m->faceGroupOf.p, m->faceGeom.p,
                                               F->V.p, F->rho.p, F->bcsDev.p, F->massFlux.p});
  flowContinuityResidual(F);
  F->hasMomAp = false;
}

// initMomentumLinearization + initAssembly + linearizeMomentum + initSolve (F/FlowModel_impl.h:522-737)
void flowAssembleMomentum(Flow* F, const fvmgpu_flow_opts& o) {
  requireReady();
  flowSyncBcs(F);
  Mesh* m = F->mesh;
  if (!(o.momentumURF > 0)) fail("flow: momentumURF must be positive");
  if (o.transient && (!F->hasVN1 || (o.time_order > 1 && !F->hasVN2))) fail("flow: transient run needs VELOCITY_N1 (and _N2)");
  parallelFor(m->nTotal, VelGradRows{m->nSelf, m->nInteriorFaces, m->row.p, m->col.p, m->entryFace.p, m->faceGroupOf.p,
                                     m->groupKindDev.p, m->faceGeom.p, F->V.p, m->gradW.p, m->nnz, F->vGrad.p});
  parallelFor(m->nTotal, PressGradRows{m->nSelf, m->nInteriorFaces, m->row.p, m->col.p, m->entryFace.p,
                                       m->faceGroupOf.p, m->groupKindDev.p, m->faceGeom.p, m->cellGeom.p, F->pFace.p,
                                       F->pGrad.p});
  flowExchange(F, F->vGrad, 9);   // GradientModel::compute ends with a sync of the gradient field
  flowExchange(F, F->pGrad, 3);   // _flowFields.pressureGradient.syncLocal(), F/FlowModel_impl.h:1011
  MomParams P;
  P.nSelf = m->nSelf; P.nInteriorFaces = m->nInteriorFaces;
  P.row = m->row.p; P.col = m->col.p; P.entryFace = m->entryFace.p; P.faceGroupOf = m->faceGroupOf.p;
  P.cellGeom = m->cellGeom.p; P.faceGeom = m->faceGeom.p;
  P.V = F->V.p; P.vGrad = F->vGrad.p; P.mu = F->mu.p; P.rho = F->rho.p; P.massFlux = F->massFlux.p;
  P.contResid = F->contResid.p; P.pGrad = F->pGrad.p; P.VN1 = F->VN1.p; P.VN2 = F->VN2.p;
  P.bcs = F->bcsDev.p;
  P.Vnew = F->Vprev.p;
  P.diag = F->mDiag.p; P.off = F->mOff.p; P.b = F->mB.p; P.isBoundary = F->comp->isBoundary.p;
  P.urf = o.momentumURF; P.timeOrder = o.transient ? o.time_order : 0; P.dt = o.dt;
  P.p = F->p.p; P.bFaceCen = m->bFaceCen.p;
  P.opPressure = o.operatingPressure; P.opTemperature = o.operatingTemperature; P.molWt = o.molecularWeight;
  P.incompressible = o.incompressible;
  parallelFor(m->nTotal, MomentumRows{P});
  // x[c1] = wall velocity for the Dirichlet ghosts; Vprev is the reference's _previousVelocity snapshot
  copyD2D(F->V.p, F->Vprev.p, 3 * (size_t)m->nTotal * sizeof(double));
  F->mDelta.zero();
}

void flowDownloadMomentum(Flow* F, double* diag3, double* off, double* b3) {
  requireReady();
  if (diag3) F->mDiag.download(diag3, 3 * (size_t)F->nTotal);
  if (off) F->mOff.download(off, (size_t)F->nnz);
  if (b3) F->mB.download(b3, 3 * (size_t)F->nTotal);
}

// LinearSolver::solve on the momentum system + postSolve + updateSolution + momAp (F/FlowModel_impl.h:744-768).
// useBcgstab: BCGStab preconditioned by `solver` with its own iteration limit / tolerances.
//
// The reference solves ONE system whose unknowns are Vector<T,3>. Norms are Vectors of per-component sums
// (F/Vector.h:189-199) while BCGStab's / CG's dot products are additionally summed over the components (reduceSum,
// F/MultiFieldReduction.cpp:165-185: shared alpha / beta / omega, see Amg::bcgstabMulti). The stationary AMG cycles
// act on every component separately -- but the convergence test is shared: `normRatio < relativeTolerance` compares the MAGNITUDE of
// the vector of component ratios (Vector::operator<, F/Vector.h:169-172; zero components divide safely and
// count 0). All components therefore take the same number of cycles / iterations N, the first one at which
// |(r_x, r_y, r_z)|_2 < tol * |(r0_x, r0_y, r0_z)|_2 (r_k = 1-norm of component k's residual). Here the components
// are advanced in rounds -- an AMG cycle continues from the stored delta, a Krylov recurrence is re-run for
// exactly N iterations -- and the shared test ends the rounds. With the loose inner tolerances of the reference's own scripts (1e-1) this is what makes the
// outer SIMPLE iterates follow the reference's.
struct SplitComponentKeep {  // scalar system of velocity component k, continuing from its stored delta
  int k; const double* diag3; const double* b3; const double* delta3; double* diag; double* b; double* delta;
  FVM_DEV void operator()(long long i) const { diag[i] = diag3[3 * i + k]; b[i] = b3[3 * i + k]; delta[i] = delta3[3 * i + k]; }
};
struct DiagTensorDiffRows {  // rows whose three diagonal components are not all equal
  const double* d3;
  FVM_DEV void operator()(long long i, double* o) const { o[0] = (d3[3 * i] != d3[3 * i + 1] || d3[3 * i] != d3[3 * i + 2]) ? 1.0 : 0.0; }
};
struct StoreComponent { int k; const double* delta; double* delta3; FVM_DEV void operator()(long long i) const { delta3[3 * i + k] = delta[i]; } };
struct ZeroComponent { int k; double* delta3; FVM_DEV void operator()(long long i) const { delta3[3 * i + k] = 0.0; } };
struct AddRows { const double* a; double* b; FVM_DEV void operator()(long long i) const { b[i] += a[i]; } };

void flowSolveMomentum(Flow* F, Amg* solver, int useBcgstab, int bcgMaxIter, double bcgRel, double bcgAbs,
                       double* rnorm0 /*3*/, int* iters /*3*/) {
  requireReady();
  Mesh* m = F->mesh;
  const int nt = m->nTotal;
  System* s = F->comp.get();
  // the reference returns the 1-norm of b per component (F/AMG.cpp:235)
  reduceRows<3>(m->nSelf, AbsRows3{F->mB.p}, F->scal.p + 2);
  if (F->multi) commAllreduceSum(F->scal.p + 2, 3);
  double norms[3];
  copyD2H(norms, F->scal.p + 2, sizeof(norms));
  copyD2D(s->off.p, F->mOff.p, (size_t)m->nnz * sizeof(double));
  bool haveHierarchy = false;
  auto select = [&](int k) {   // component k (with its current delta) becomes the scalar system
    parallelFor(nt, SplitComponentKeep{k, F->mDiag.p, F->mB.p, F->mDelta.p, s->diag.p, s->b.p, s->delta.p});
    bool same = false;
    if (haveHierarchy) {
      reduceRows<1>(m->nSelf, DiffCountRows{s->diag.p, F->lastDiag.p}, F->scal.p + 5);
      if (F->multi) commAllreduceSum(F->scal.p + 5, 1);  // every rank must take the same decision
      double nd;
      copyD2H(&nd, F->scal.p + 5, sizeof(double));
      same = nd == 0.0;
    }
    if (!same) {
      s->version = nextVersion();  // the hierarchy is rebuilt for this component's diagonal
      copyD2D(F->lastDiag.p, s->diag.p, (size_t)nt * sizeof(double));
      haveHierarchy = true;
    }
  };
  auto store = [&](int k) { parallelFor(nt, StoreComponent{k, s->delta.p, F->mDelta.p}); };
  auto krylov = [&](int maxIter, double rel, double abs, int* it) {   // 1: one AMG cycle, 2: the reference's ILU(0)
    const int keep = solver->precondKind;
    double r0 = 0, r = 0;
    solver->precondKind = useBcgstab == 2 ? 1 : 0;
    solver->bcgstab(s, maxIter, rel, abs, &r0, &r, it);
    solver->precondKind = keep;
  };
  const fvmgpu_amg_opts userOpts = solver->opts;
  const bool jacobi = useBcgstab == 3;   // JacobiSolver (F/JacobiSolver.cpp:46-95) with the limits passed like BCGStab's
  if (jacobi) useBcgstab = 0;
  const int limit = useBcgstab ? bcgMaxIter : (jacobi ? bcgMaxIter - 1 : userOpts.nMaxIterations - 1);
  const double rel = (useBcgstab || jacobi) ? bcgRel : userOpts.relativeTolerance;
  const double abs = (useBcgstab || jacobi) ? bcgAbs : userOpts.absoluteTolerance;
  std::vector<double> hist[3];
  for (int k = 0; k < 3; k++) hist[k].assign(1, norms[k]);
  // ---- the shared convergence test: |rNorm|_2 / |rNorm0|_2 over the vector of component 1-norms
  // (MultiFieldReduction::normalize + Vector::operator<, F/AMG.cpp:256-272, F/BCGStab.cpp:131-147)
  double den = 0;
  for (int k = 0; k < 3; k++) den += norms[k] * norms[k];
  auto converged = [&](int i) {
    double num = 0;
    double ratios = 0;   // JacobiSolver divides component by component (safeDivide, F/JacobiSolver.cpp:73)
    for (int k = 0; k < 3; k++) {
      const double r = hist[k][(size_t)std::min<int>(i, (int)hist[k].size() - 1)];
      num += r * r;
      const double q = norms[k] != 0.0 ? r / norms[k] : r;
      ratios += q * q;
    }
    if (jacobi) return num < abs * abs || ratios < rel * rel;
    return num < abs * abs || (den > 0 ? num / den : num) < rel * rel;
  };
  int N = 0;
  bool coupledKrylov = false;
  // one diagonal for all three components (always the case without symmetry planes)? Then the system is ONE matrix
  // with Vector<T,3> unknowns: the reference's BCGStab shares its scalars between the components
  // (Amg::bcgstabMulti), and the AMG cycles run once for all components (Amg::solveMulti)
  bool oneDiagonal = false;
  const bool wantMultiRhs = !useBcgstab && !jacobi && solver->multiRhsSupported();
  if ((useBcgstab || wantMultiRhs) && den > 0 && !(den < abs * abs)) {
    reduceRows<1>(m->nSelf, DiagTensorDiffRows{F->mDiag.p}, F->scal.p + 5);
    if (F->multi) commAllreduceSum(F->scal.p + 5, 1);
    double nd;
    copyD2H(&nd, F->scal.p + 5, sizeof(double));
    oneDiagonal = nd == 0.0;
    coupledKrylov = useBcgstab && oneDiagonal;
  }
  if (useBcgstab == 4 && !coupledKrylov && den > 0 && !(den < abs * abs))
    fail("flow: CG on the momentum system needs one diagonal for the three components (no symmetry planes)");
  if (coupledKrylov) {
    select(0);
    const int keep = solver->precondKind;
    solver->precondKind = useBcgstab == 2 ? 1 : 0;
    double r0v[3], rv[3];
    if (useBcgstab == 4) solver->cgMulti(s, 3, F->mB.p, F->mDelta.p, bcgMaxIter, bcgRel, bcgAbs, r0v, rv, &N);
    else solver->bcgstabMulti(s, 3, F->mB.p, F->mDelta.p, bcgMaxIter, bcgRel, bcgAbs, r0v, rv, &N);
    solver->precondKind = keep;
  } else if (wantMultiRhs && oneDiagonal && !F->multi) {
    // every matrix entry is read once per pass for all components; a 2-D flow carries two of them
    select(0);
    const int nc = (m->dim == 2 && !(norms[2] > 0.0)) ? 2 : 3;
    double r0v[3] = {0, 0, 0}, rv[3] = {0, 0, 0};
    solver->solveMulti(s, nc, F->mB.p, F->mDelta.p, 3, limit + 1, rel, abs, r0v, rv, &N);
  } else if (den > 0 && !(den < abs * abs)) {
    while (N < limit) {
      N++;
      for (int k = 0; k < 3; k++) {
        if (!(norms[k] > 0.0)) continue;
        if (useBcgstab) {   // a Krylov recurrence cannot be resumed: run it again for exactly N iterations
          parallelFor(nt, ZeroComponent{k, F->mDelta.p});
          select(k);
          int done = 0;
          krylov(N, 0.0, 0.0, &done);
          hist[k] = solver->history;
        } else {            // stationary iteration: one more cycle from the stored delta
          select(k);
          double r0 = 0, r = 0;
          int one = 0;
          if (jacobi) {
            solver->jacobiSolve(s, 2, 0.0, 0.0, &r0, &r, &one);
          } else {
            solver->opts.nMaxIterations = 2; solver->opts.relativeTolerance = 0.0; solver->opts.absoluteTolerance = 0.0;
            solver->solve(s, &r0, &r, &one);
            solver->opts = userOpts;
          }
          hist[k].push_back(r);
        }
        store(k);
      }
      if (converged(N)) break;
    }
  }
  for (int k = 0; k < 3; k++) {
    if (rnorm0) rnorm0[k] = norms[k];
    if (iters) iters[k] = (norms[k] > 0.0 || coupledKrylov) ? N : 0;   // (a zero component stays zero whichever path ran)
  }
  parallelFor(3LL * nt, AddRows{F->mDelta.p, F->V.p});   // ls.updateSolution: x += delta
  parallelFor(nt - m->nSelf, MomentumGhostRows{m->nSelf, m->row.p, m->col.p, s->isBoundary.p, F->mDiag.p, F->mOff.p,
                                               F->mB.p, F->mDelta.p, F->V.p});
  copyD2D(F->momAp.p, F->mDiag.p, 3 * (size_t)nt * sizeof(double));
  flowExchange(F, F->momAp, 3);   // _momApField->syncLocal(), F/FlowModel_impl.h:768
  F->hasMomAp = true;
}

// initContinuityLinearization + initAssembly + linearizeContinuity + initSolve (F/FlowModel_impl.h:998-1198,1394-1407)
void flowAssembleContinuity(Flow* F, const fvmgpu_flow_opts& o) {
  requireReady();
  flowSyncBcs(F);
  if (!F->hasMomAp) fail("flow: continuity needs the momentum coefficients (solve the momentum equations first)");
  Mesh* m = F->mesh;
  ContParams P = contParams(F, o.momentumURF);
  parallelFor(m->nFaces, MassFluxFaces{P});
  const int nb = m->nFaces - m->nInteriorFaces;
  reduceRows<1>(nb, BoundaryFluxSum{m->nInteriorFaces, F->massFlux.p, F->bcsDev.p, m->faceGroupOf.p}, F->scal.p + 0);
  reduceRows<1>(m->nSelf, VolumeSum{m->cellGeom.p}, F->scal.p + 1);
  if (F->multi) commAllreduceSum(F->scal.p, 2);   // F/FlowModel_impl.h:1134-1141, 1167-1169
  System* s = F->pp.get();
  // a pressure boundary anchors the pressure level: no reference cell, no net-flux redistribution (:1052-1056)
  ContinuityRows K{P, F->scal.p, F->pressureBoundaryAnywhere ? 0 : 1, F->refCell, s->diag.p, s->off.p, s->b.p, s->isBoundary.p};
  parallelFor(m->nTotal, K);
  s->delta.zero();
  s->version = nextVersion();
}

void flowDownloadContinuity(Flow* F, double* diag, double* off, double* b, int* isBoundary) {
  requireReady();
  System* s = F->pp.get();
  if (diag) s->diag.download(diag, (size_t)F->nTotal);
  if (off) s->off.download(off, (size_t)F->nnz);
  if (b) s->b.download(b, (size_t)F->nTotal);
  if (isBoundary) s->isBoundary.download(isBoundary, (size_t)F->nTotal);
}

// solve + postSolve + postContinuitySolve (F/FlowModel_impl.h:1410-1430, 1263-1339)
void flowSolveContinuity(Flow* F, Amg* solver, int useBcgstab, int bcgMaxIter, double bcgRel, double bcgAbs,
                         const fvmgpu_flow_opts& o, double* rnorm0, int* iters) {
  requireReady();
  Mesh* m = F->mesh;
  System* s = F->pp.get();
  double r0 = 0, r = 0;
  int it = 0;
  if (useBcgstab == 3) {
    solver->jacobiSolve(s, bcgMaxIter, bcgRel, bcgAbs, &r0, &r, &it);
  } else if (useBcgstab == 4) {
    solver->cg(s, bcgMaxIter, bcgRel, bcgAbs, &r0, &r, &it);
  } else if (useBcgstab) {
    const int keep = solver->precondKind;
    solver->precondKind = useBcgstab == 2 ? 1 : 0;
    solver->bcgstab(s, bcgMaxIter, bcgRel, bcgAbs, &r0, &r, &it);
    solver->precondKind = keep;
  } else solver->solve(s, &r0, &r, &it);
  if (rnorm0) *rnorm0 = r0;
  if (iters) *iters = it;
  double* pp = s->delta.p;
  parallelFor(m->nTotal - m->nSelf, PpGhostRows{m->nSelf, m->row.p, m->col.p, s->diag.p, s->off.p, s->b.p, pp});
  if (F->multi) m->halo.exchange(pp, 1);  // the interface ghosts hold the owner's pp again (their rows are not equations here)
  // setReferencePP: pp of the reference cell (read on the device, no host round trip); across ranks the
  // owner of the globally lowest cell contributes it and the others 0 (F/FlowModel_impl.h:1216-1230)
  const double* refPP = nullptr;
  if (!F->pressureBoundaryAnywhere) {
    if (F->multi) {
      parallelFor(1, ReferencePpKernel{pp, F->refCell, F->scal.p + 6});
      commAllreduceSum(F->scal.p + 6, 1);
      refPP = F->scal.p + 6;
    } else {
      refPP = pp + F->refCell;
    }
  }
  parallelFor(m->nTotal, CorrectPressureRows{pp, refPP, o.pressureURF, F->p.p});
  parallelFor(m->nInteriorFaces, CorrectMassFluxFaces{m->faceCells.p, m->pairToCol.p, s->off.p, pp, F->massFlux.p});
  if (F->multi)
    parallelFor(m->nFaces - m->nInteriorFaces, CorrectInterfaceMassFluxFaces{m->nInteriorFaces, m->faceCells.p, F->bcsDev.p,
                                                                             m->faceGroupOf.p, F->pCoeff.p, pp, F->massFlux.p});
  if (F->hasPressureBoundary)  // correctMassFluxBoundary: massFlux -= dMassFlux, the flux rows' solution (:831-842)
    parallelFor(m->nFaces - m->nInteriorFaces, CorrectBoundaryMassFluxFaces{m->nInteriorFaces, m->faceCells.p, F->bCoeffL.p,
                                                                            pp, F->massFlux.p});
  ContParams P = contParams(F, o.momentumURF);
  if (o.correctVelocity) {
    parallelFor(m->nSelf, CorrectVelocityRows{P, pp, F->V.p});  // in place: a row reads only its own velocity
  }
  if (F->hasPressureBoundary)
    parallelFor(m->nFaces - m->nInteriorFaces, PressureBoundaryPostFaces{P, pp, F->V.p, F->p.p});
  flowExchange(F, F->V, 3);   // _flowFields.velocity.syncLocal(); pressure.syncLocal(), F/FlowModel_impl.h:1334-1335
  flowExchange(F, F->p, 1);
  parallelFor(m->nFaces, FacePressureFaces{P, F->pFace.p});
  flowContinuityResidual(F);
  F->hasMomAp = false;  // the reference discards momAp after the continuity step
}

// F/FlowModel_impl.h:931-994: the reference cell is the globally lowest fluid cell; the rank that owns it
// passes its local index, every other rank -1
void flowSetReferenceCell(Flow* F, int localCell) {
  if (localCell >= F->nSelf) fail("flow_set_reference_cell: cell %d is not an interior cell of this mesh part", localCell);
  F->refCell = localCell;
}

void flowDestroy(Flow* F) { delete F; }

}  // namespace fvmgpu
