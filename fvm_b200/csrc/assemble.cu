// fvm_b200 / libfvmgpu -- finite-volume assembly kernels (FP64, sm_100a).
//
//   k_gradient       GradientMatrix::getGradient + boundary copy/reflect
//                    (F/GradientMatrix.h:55-76, F/GradientModel.h:530-566)
//   k_assemble       ONE fused, row-parallel gather kernel that performs, per matrix row and in
//                    the reference's own order of floating-point operations,
//                      DiffusionDiscretization   F/DiffusionDiscretization.h:165-228
//                      ConvectionDiscretization  F/ConvectionDiscretization.h:119-199
//                      SourceDiscretization      F/SourceDiscretization.h:54-57
//                      TimeDerivativeDiscr.      F/TimeDerivativeDiscretization.h:102-108,149-155
//                      GenericBCS::apply*BC      F/GenericBCS.h:77-356
//                      Underrelaxer              F/Underrelaxer.h:49-52
//                      eliminateBoundaryEquations F/CRMatrix.h:899-944,1064-1085
//   k_post_solve     LinearSystem::postSolve + updateSolution (F/LinearSystem.cpp:250-269,
//                    F/CRMatrix.h:433-454, F/FluxJacobianMatrix.h, F/DiagonalMatrix.h)
//
// The reference scatters face contributions into diag[c0], diag[c1], r[c0], r[c1] in global
// face order. Here each row GATHERS its faces through entryFace[] in the same (ascending face)
// order, so every output is written exactly once, no atomics are needed and the sums are
// bit-reproducible. This file is compiled with -fmad=false: the reference's x86-64 build has no
// FMA contraction, so IEEE add/mul/div in the same order give identical bits.
#include "structs.cuh"

namespace fvmgpu {

struct AsmParams {
  int nSelf, nTotal, nInteriorFaces, nGroups;
  const int* row;
  const int* col;
  const int* entryFace;
  const int* faceGroupOf;
  const double4* cellGeom;
  const double4* faceGeom;
  const double4* cellState;  // {gx,gy,gz,x}
  const double* diffusivity;
  const double* source;
  const double* faceFlux;
  const double* xN1;
  const double* xN2;
  const double* density;
  const double* contResid;
  const BcEntry* bcs;
  double* x;  // Dirichlet ghost values are written here (assembly itself reads x from cellState)
  double* diag;
  double* off;
  double* b;
  int* isBoundary;
  double* bflux;
  double* rflux;
  double* coeffL;
  double* coeffR;
  fvmgpu_assemble_opts o;
};

// ---------------------------------------------------------------- gradient
FVM_DEV void rowGradient(int i, double xi, const int* row,
                                            const int* col, const double* x,
                                            const double* w, long long nnz, double& g0,
                                            double& g1, double& g2) {
  g0 = 0.0; g1 = 0.0; g2 = 0.0;
  const int r0 = row[i], r1 = row[i + 1];
  for (int k = r0; k < r1; k++) {
    const double v = x[col[k]] - xi;  // Gradient::accumulate(wt, x[j]-x[nr])  F/Gradient.h:57-61
    g0 += w[k] * v;
    g1 += w[nnz + k] * v;
    g2 += w[2 * nnz + k] * v;
  }
}

struct GradientRows {
  int nSelf, nInteriorFaces; const int* row; const int* col; const int* entryFace; const int* faceGroupOf;
  const int* groupKind; const double4* faceGeom; const double* x; const double* w; long long nnz;
  double4* state;
  FVM_DEV void operator()(long long ii) const {
  const int i = (int)ii;
  const double xi = x[i];
  double g0, g1, g2;
  if (i < nSelf) {
    rowGradient(i, xi, row, col, x, w, nnz, g0, g1, g2);
    state[i] = make_double4(g0, g1, g2, xi);
    return;
  }
  // ghost cell: one face. Boundary groups copy the neighbour's gradient, symmetry groups reflect
  // it; interface ghosts are filled by the halo exchange (only x is refreshed here).
  const int k = row[i];
  if (row[i + 1] - k != 1) { state[i] = make_double4(0, 0, 0, xi); return; }
  const int f = entryFace[k] >> 1;
  const int c0 = col[k];
  const int kind = (f >= nInteriorFaces) ? groupKind[faceGroupOf[f - nInteriorFaces]] : FVMGPU_GROUP_INTERIOR;
  if (kind == FVMGPU_GROUP_INTERFACE || kind == FVMGPU_GROUP_INTERIOR) {
    double4 s = state[i];
    s.w = xi;
    state[i] = s;
    return;
  }
  if (kind == FVMGPU_GROUP_DIELECTRIC_INTERFACE) {
    // no copy from the neighbour (F/GradientModel.h:538-539) and GradientMatrix::getGradient only fills the interior
    // rows (F/GradientMatrix.h:55-76): the reference leaves the gradient of these ghost cells unset -- here it is 0
    state[i] = make_double4(0.0, 0.0, 0.0, xi);
    return;
  }
  rowGradient(c0, x[c0], row, col, x, w, nnz, g0, g1, g2);
  if (kind == FVMGPU_GROUP_SYMMETRY) {  // reflectGradient, F/GradientModel.h:21-28
    const double4 fg = faceGeom[f];
    const double e0 = fg.x / fg.w, e1 = fg.y / fg.w, e2 = fg.z / fg.w;
    double dot = 0.0;
    dot += g0 * e0; dot += g1 * e1; dot += g2 * e2;
    const double t = 2.0 * dot;
    g0 = g0 - t * e0; g1 = g1 - t * e1; g2 = g2 - t * e2;
  }
  state[i] = make_double4(g0, g1, g2, xi);
  }
};

// ---------------------------------------------------------------- face physics
struct CellV {
  double4 g;  // cx,cy,cz,vol
  double4 s;  // gx,gy,gz,x
  double k;   // diffusivity
};

FVM_DEV CellV loadCell(const AsmParams& P, int c) {
  CellV v;
  v.g = P.cellGeom[c];
  v.s = P.cellState[c];
  v.k = P.diffusivity ? P.diffusivity[c] : 1.0;
  return v;
}

// harmonicAverage, F/DiffusionDiscretization.h:19-27
FVM_DEV double harmonicAverage(double x0, double x1) {
  const double sum = x0 + x1;
  if (x0 + x1 != 0.0) return 2.0 * x0 * x1 / sum;
  return sum;
}

// F/DiffusionDiscretization.h:165-209 for one face with cells (c0,c1)
FVM_DEV void diffusionFaceRegular(const double4 fg, const CellV& a0, const CellV& a1,
                                              double& diffCoeff, double& dFlux) {
  const double vol0 = a0.g.w, vol1 = a1.g.w;
  const double ds0 = a1.g.x - a0.g.x, ds1 = a1.g.y - a0.g.y, ds2 = a1.g.z - a0.g.z;
  double fd;
  if (vol0 == 0.) fd = a1.k;
  else if (vol1 == 0.) fd = a0.k;
  else fd = harmonicAverage(a0.k, a1.k);
  const double diffMetric = fg.w * fg.w / (fg.x * ds0 + fg.y * ds1 + fg.z * ds2);
  diffCoeff = fd * diffMetric;
  const double sc0 = fd * (fg.x - ds0 * diffMetric);
  const double sc1 = fd * (fg.y - ds1 * diffMetric);
  const double sc2 = fd * (fg.z - ds2 * diffMetric);
  const double vs = vol0 + vol1;
  const double gf0 = (a0.s.x * vol0 + a1.s.x * vol1) / vs;
  const double gf1 = (a0.s.y * vol0 + a1.s.y * vol1) / vs;
  const double gf2 = (a0.s.z * vol0 + a1.s.z * vol1) / vs;
  double sec = 0.0;  // Gradient * Vector, F/Gradient.h:99-105
  sec += gf0 * sc0; sec += gf1 * sc1; sec += gf2 * sc2;
  dFlux = diffCoeff * (a1.s.w - a0.s.w) + sec;
}

// The "dielectric interface" branch of the same loop (F/DiffusionDiscretization.h:97-151): on the faces of a group of
// that type the two cells are separated by a thin layer of the given thickness -- the metric is
// sign(A . ds) |A| / (|ds| + thickness / 2), the face diffusivity always the harmonic mean, no secondary-gradient term.
FVM_DEV void diffusionFaceDielectric(const double4 fg, const CellV& a0, const CellV& a1, double thickness,
                                     double& diffCoeff, double& dFlux) {
  const double ds0 = a1.g.x - a0.g.x, ds1 = a1.g.y - a0.g.y, ds2 = a1.g.z - a0.g.z;
  double m2 = 0.0;   // mag(ds) = sqrt(dot(ds, ds)), F/Vector.h
  m2 += ds0 * ds0; m2 += ds1 * ds1; m2 += ds2 * ds2;
  const double dsMag = sqrt(m2);
  const double fd = harmonicAverage(a0.k, a1.k);
  double sign = 1.0;
  double ad = 0.0;
  ad += fg.x * ds0; ad += fg.y * ds1; ad += fg.z * ds2;
  if (ad < 0.0) sign *= -1.0;
  const double diffMetric = sign * fg.w / (dsMag + 0.5 * thickness);
  diffCoeff = fd * diffMetric;
  dFlux = diffCoeff * (a1.s.w - a0.s.w);
}
FVM_DEV void diffusionFace(const AsmParams& P, int f, const double4 fg, const CellV& a0, const CellV& a1,
                           double& diffCoeff, double& dFlux) {
  if (f >= P.nInteriorFaces && P.bcs[P.faceGroupOf[f - P.nInteriorFaces]].groupKind == FVMGPU_GROUP_DIELECTRIC_INTERFACE)
    diffusionFaceDielectric(fg, a0, a1, P.o.interface_thickness, diffCoeff, dFlux);
  else
    diffusionFaceRegular(fg, a0, a1, diffCoeff, dFlux);
}

// Values of the (single-face) ghost row c1 and of the interior coefficient toward it, as the
// reference leaves them after the discretization list and before the BC loop.
struct GhostRow {
  double r1, diag1, c10;
};

FVM_DEV GhostRow ghostRowBeforeBC(const AsmParams& P, int f, const double4 fg,
                                                     const CellV& a0, const CellV& a1) {
  GhostRow g;
  g.r1 = 0.0; g.diag1 = 0.0; g.c10 = 0.0;
  if (P.o.diffusion) {
    double dc, df;
    diffusionFace(P, f, fg, a0, a1, dc, df);
    g.r1 -= df;
    g.c10 += dc;
    g.diag1 -= dc;
  }
  if (P.o.convection) {
    const double flux = P.faceFlux[f];
    double varFlux;
    if (P.o.convection == 2) varFlux = 0.5 * flux * (a0.s.w + a0.s.w);  // reference quirk :131
    else varFlux = (flux > 0.0) ? flux * a0.s.w : flux * a1.s.w;
    if (flux > 0.0) g.c10 += flux;
    else g.diag1 += flux;
    g.r1 += varFlux;
  }
  return g;
}

struct BcOut {
  bool marks;      // CRMatrix::setBoundary(c1)
  bool setsX;      // Dirichlet writes x[c1]
  double xNew;
  double flux, rflux, cL, cR;
};

// GenericBCS::apply*BC restricted to what they do to the ghost row (r1, diag1, c10) and to the
// boundary-flux side system. c01 is the interior row's coefficient toward the ghost.
FVM_DEV BcOut bcOnGhost(int kind, const double* p, double bValue, double areaMag,
                                           double x0, double x1, double c01, GhostRow& g) {
  BcOut o;
  o.marks = false; o.setsX = false; o.xNew = x1; o.flux = 0; o.rflux = 0; o.cL = 0; o.cR = 0;
  const double sb = 5.670373E-8;
  switch (kind) {
    case FVMGPU_BC_DIRICHLET: {  // :77-115
      const double fluxB = -g.r1;
      const double dFluxdXC0 = -g.c10;
      const double dFluxdXC1 = -g.diag1;
      const double dXC1 = bValue - x1;
      const double dFlux = dFluxdXC1 * dXC1;
      o.setsX = true; o.xNew = bValue;
      g.c10 = 0.0; g.r1 = 0.0; g.diag1 = -1.0;
      o.cL = dFluxdXC0; o.cR = 0.0; o.flux = fluxB; o.rflux = dFlux;
    } break;
    case FVMGPU_BC_NEUMANN: {  // :129-157
      const double fluxB = -g.r1;
      const double dFlux = bValue * areaMag - fluxB;
      g.r1 = dFlux;
      o.marks = true;
      o.flux = bValue * areaMag;
    } break;
    case FVMGPU_BC_EXTRAPOLATION: {  // :180-212
      const double fluxB = -g.r1;
      const double dFluxdXC0 = -g.c10;
      const double xc0mxc1 = x0 - x1;
      g.diag1 = -1.0; g.c10 = 1.0; g.r1 = xc0mxc1;
      o.marks = true;
      o.cL = dFluxdXC0; o.cR = dFluxdXC0; o.flux = fluxB; o.rflux = 0.0;
    } break;
    case FVMGPU_BC_CONVECTIVE: {  // :214-245  p0 = h, p1 = Xinf
      const double h = bValue, Xinf = p[1];
      const double fluxInterior = -g.r1;
      const double fluxBoundary = -h * (x1 - Xinf) * areaMag;
      g.r1 = fluxBoundary - fluxInterior;
      g.diag1 -= h * areaMag;
      o.marks = true;
      o.flux = fluxBoundary; o.rflux = 0.0; o.cL = 0.0; o.cR = -h * areaMag;
    } break;
    case FVMGPU_BC_DIELECTRIC_INTERFACE: {  // applyDielectricInterfaceBC :367-400  p0 = Xinf, p1 = hCoeff, p2 = source
      const double Xinf = bValue, h = p[1], source = p[2];
      const double fluxInterior = -g.r1;
      double fluxSource = source * areaMag;
      fluxSource /= 2.0;   // (the reference halves it in 2-D and in 3-D alike)
      const double fluxBoundary = -h * (x1 - Xinf) * areaMag + fluxSource;
      g.r1 = fluxBoundary - fluxInterior;
      g.diag1 -= h * areaMag;
      o.marks = true;
      o.flux = fluxBoundary; o.rflux = 0.0; o.cL = 0.0; o.cR = -h * areaMag;
    } break;
    case FVMGPU_BC_RADIATIVE: {  // :253-288  p0 = emissivity, p1 = Xinf
      const double em = bValue, Xinf = p[1];
      const double fluxInterior = -g.r1;
      const double fluxBoundary = -em * sb * (x1 * x1 * x1 * x1 - Xinf * Xinf * Xinf * Xinf) * areaMag;
      g.r1 = fluxBoundary - fluxInterior;
      g.diag1 -= 4 * em * sb * x1 * x1 * x1 * areaMag;
      o.marks = true;
      o.flux = fluxBoundary; o.rflux = 0.0; o.cL = 0.0; o.cR = -4 * em * sb * x1 * x1 * x1 * areaMag;
    } break;
    case FVMGPU_BC_MIXED: {  // :290-323  p0 = h, p1 = emissivity, p2 = Xinf
      const double h = bValue, em = p[1], Xinf = p[2];
      const double fluxInterior = -g.r1;
      const double fluxBoundary =
          (-em * sb * (x1 * x1 * x1 * x1 - Xinf * Xinf * Xinf * Xinf) - h * (x1 - Xinf)) * areaMag;
      g.r1 = fluxBoundary - fluxInterior;
      g.diag1 -= (4 * em * sb * x1 * x1 * x1 + h) * areaMag;
      o.marks = true;
      o.flux = fluxBoundary; o.rflux = 0.0; o.cL = 0.0; o.cR = -4 * em * sb * x1 * x1 * x1 * areaMag;
    } break;
    case FVMGPU_BC_INTERFACE: {  // :325-356 (ghost is c1, sign = +1)
      const double fluxInterior = -g.r1;
      o.cL = -1.0 * g.c10; o.cR = 1.0 * c01;
      g.r1 = 0.0; g.c10 = 0.0;
      o.flux = fluxInterior; o.rflux = 0.0;
    } break;
    default: break;
  }
  return o;
}

FVM_DEV int effectiveBcKind(const BcEntry& bc, const AsmParams& P, int f) {
  int kind = bc.kind;
  if (kind == FVMGPU_BC_DIRICHLET_OR_OUTFLOW) {  // F/ThermalModel_impl.h:313-331
    if (P.faceFlux && P.faceFlux[f] > 0.) kind = FVMGPU_BC_EXTRAPOLATION;
    else kind = FVMGPU_BC_DIRICHLET;
  }
  return kind;
}

#define MAX_BC_GROUPS 64

struct AssembleRows {
  AsmParams P;
  FVM_DEV void operator()(long long ii) const {
  const int i = (int)ii;
  const BcEntry* sbc = P.bcs;  // a handful of 64-byte entries: stays in L1 / constant-like reads
  const int r0 = P.row[i], r1 = P.row[i + 1];
  const CellV me = loadCell(P, i);
  double diag = 0.0, r = 0.0;
  bool hasB = false;

  // ---- DiffusionDiscretization (all face groups in face order)
  for (int k = r0; k < r1; k++) {
    const int ef = P.entryFace[k];
    const int f = ef >> 1, side = ef & 1;
    if (f >= P.nInteriorFaces) hasB = true;
    double offk = 0.0;
    if (P.o.diffusion) {
      const CellV ot = loadCell(P, P.col[k]);
      const double4 fg = P.faceGeom[f];
      double dc, df;
      if (side == 0) { diffusionFace(P, f, fg, me, ot, dc, df); r += df; }
      else { diffusionFace(P, f, fg, ot, me, dc, df); r -= df; }
      offk += dc;
      diag -= dc;
    }
    P.off[k] = offk;
  }
  // ---- ConvectionDiscretization
  if (P.o.convection) {
    for (int k = r0; k < r1; k++) {
      const int ef = P.entryFace[k];
      const int f = ef >> 1, side = ef & 1;
      const double flux = P.faceFlux[f];
      const double xo = P.cellState[P.col[k]].w;
      const double x0 = side ? xo : me.s.w, x1 = side ? me.s.w : xo;
      double varFlux;
      if (P.o.convection == 2) varFlux = 0.5 * flux * (x0 + x0);
      else varFlux = (flux > 0.0) ? flux * x0 : flux * x1;
      if (side == 0) {  // this row is c0
        if (flux > 0.0) diag -= flux;
        else P.off[k] -= flux;  // coeff01
        r -= varFlux;
      } else {          // this row is c1
        if (flux > 0.0) P.off[k] += flux;  // coeff10
        else diag += flux;
        r += varFlux;
      }
    }
    if (i < P.nSelf && P.contResid) diag += P.contResid[i];
  }
  if (i < P.nSelf) {
    // ---- SourceDiscretization
    if (P.o.source && P.source) r += me.g.w * P.source[i];
    // ---- TimeDerivativeDiscretization (static mesh branches)
    if (P.o.time_order == 1) {
      const double rhoVbydT = P.density[i] * me.g.w / P.o.dt;
      r -= rhoVbydT * (me.s.w - P.xN1[i]);
      diag -= rhoVbydT;
    } else if (P.o.time_order == 2) {
      const double rhoVbydT = P.density[i] * me.g.w / P.o.dt;
      r -= rhoVbydT * (1.5 * me.s.w - 2.0 * P.xN1[i] + 0.5 * P.xN2[i]);
      diag -= rhoVbydT * 1.5;
    }
  }

  if (i >= P.nSelf) {
    // ================= ghost row: exactly one face =================
    int marks = 0;
    if (hasB && P.o.apply_bcs && r1 - r0 == 1) {
      const int k = r0;
      const int f = P.entryFace[k] >> 1;
      const int gi = P.faceGroupOf[f - P.nInteriorFaces];
      const BcEntry& bc = sbc[gi];
      if (bc.kind >= 0) {
        const int kind = effectiveBcKind(bc, P, f);
        const double bValue = bc.perFace ? bc.perFace[f - bc.offset] : bc.p[0];
        const int c0 = P.col[k];
        const double x0 = P.cellState[c0].w;
        GhostRow g;
        g.r1 = r; g.diag1 = diag; g.c10 = P.off[k];
        // c01 (only the interface BC reads it): the interior row's coefficient toward this ghost
        double c01 = 0.0;
        if (kind == FVMGPU_BC_INTERFACE) {
          const CellV a0 = loadCell(P, c0);
          const double4 fg = P.faceGeom[f];
          if (P.o.diffusion) { double dc, df; diffusionFace(P, f, fg, a0, me, dc, df); c01 += dc; }
          if (P.o.convection && !(P.faceFlux[f] > 0.0)) c01 -= P.faceFlux[f];
        }
        const BcOut o = bcOnGhost(kind, bc.p, bValue, P.faceGeom[f].w, x0, me.s.w, c01, g);
        r = g.r1; diag = g.diag1; P.off[k] = g.c10;
        marks = o.marks ? 1 : 0;
        if (o.setsX) P.x[i] = o.xNew;
        const int bf = f - P.nInteriorFaces;
        P.bflux[bf] = o.flux; P.rflux[bf] = o.rflux; P.coeffL[bf] = o.cL; P.coeffR[bf] = o.cR;
      }
    }
    P.isBoundary[i] = marks;
    P.diag[i] = diag;
    P.b[i] = r;
    return;
  }

  // ================= interior row =================
  if (hasB && P.o.apply_bcs) {
    // ---- BC loop: effect of each boundary face's BC on THIS row, in face order
    for (int k = r0; k < r1; k++) {
      const int f = P.entryFace[k] >> 1;
      if (f < P.nInteriorFaces) continue;
      const BcEntry& bc = sbc[P.faceGroupOf[f - P.nInteriorFaces]];
      if (bc.kind < 0) continue;
      const int kind = effectiveBcKind(bc, P, f);
      const int c1 = P.col[k];
      if (kind == FVMGPU_BC_DIRICHLET) {
        const double bValue = bc.perFace ? bc.perFace[f - bc.offset] : bc.p[0];
        const double dXC1 = bValue - P.cellState[c1].w;
        const double dRC0 = P.off[k] * dXC1;
        r += dRC0;
        P.off[k] = 0.0;
      } else if (kind == FVMGPU_BC_EXTRAPOLATION) {
        const CellV a1 = loadCell(P, c1);
        const GhostRow g = ghostRowBeforeBC(P, f, P.faceGeom[f], me, a1);
        const double dFluxdXC1 = -g.diag1;
        const double xc0mxc1 = me.s.w - a1.s.w;
        diag += dFluxdXC1;
        r += dFluxdXC1 * xc0mxc1;
        P.off[k] = 0.0;
      }
    }
  }
  // ---- Underrelaxer
  if (P.o.underrelax > 0.0) diag /= P.o.underrelax;
  // ---- LinearSystem::initSolve -> eliminateBoundaryEquations, in ghost-row (= face) order
  if (hasB && P.o.apply_bcs && P.o.eliminate_boundary) {
    for (int k = r0; k < r1; k++) {
      const int f = P.entryFace[k] >> 1;
      if (f < P.nInteriorFaces) continue;
      const BcEntry& bc = sbc[P.faceGroupOf[f - P.nInteriorFaces]];
      if (bc.kind < 0) continue;
      const int kind = effectiveBcKind(bc, P, f);
      if (kind == FVMGPU_BC_DIRICHLET || kind == FVMGPU_BC_INTERFACE) continue;  // not marked
      const int c1 = P.col[k];
      const CellV a1 = loadCell(P, c1);
      const double4 fg = P.faceGeom[f];
      GhostRow g = ghostRowBeforeBC(P, f, fg, me, a1);
      const double bValue = bc.perFace ? bc.perFace[f - bc.offset] : bc.p[0];
      const BcOut o = bcOnGhost(kind, bc.p, bValue, fg.w, me.s.w, a1.s.w, 0.0, g);
      if (!o.marks) continue;
      const double a_ij = P.off[k];
      diag -= a_ij * (g.c10 / g.diag1);
      r -= a_ij * (g.r1 / g.diag1);
      P.off[k] = 0.0;
    }
  }
  P.isBoundary[i] = 0;
  P.diag[i] = diag;
  P.b[i] = r;
  }
};

// ---------------------------------------------------------------- post solve
struct PostSolveRows {
  int nSelf, nInteriorFaces; const int* row; const int* col; const int* entryFace; const double* diag;
  const double* off; const double* b; const int* isBoundary; double* delta; double* x; double* bflux;
  const double* rflux; const double* coeffL; const double* coeffR; int hasFluxRows;
  FVM_DEV void operator()(long long ii) const {
  const int i = (int)ii;
  if (i < nSelf) {
    x[i] += delta[i];
    return;
  }
  // ghost row: CRMatrix::solveBoundary (F/CRMatrix.h:433-454) then the flux row
  double dj = delta[i];
  const int r0 = row[i], r1 = row[i + 1];
  if (isBoundary[i]) {
    double sum = b[i];
    for (int k = r0; k < r1; k++) sum += off[k] * delta[col[k]];
    dj = -sum / diag[i];
    delta[i] = dj;
  }
  x[i] += dj;
  if (hasFluxRows && r1 - r0 == 1) {
    const int f = entryFace[r0] >> 1;
    if (f >= nInteriorFaces) {
      const int bf = f - nInteriorFaces;
      double rr = rflux[bf];
      rr += coeffL[bf] * delta[col[r0]] + coeffR[bf] * dj;  // FluxJacobianMatrix::multiplyAndAdd
      const double dflux = -rr / -1.0;                      // DiagonalMatrix (dFluxdFlux = -1) forwardGS
      bflux[bf] += dflux;
    }
  }
  }
};

struct FillKernel {
  double* p; double v;
  FVM_DEV void operator()(long long i) const { p[i] = v; }
};
struct GradToAosKernel {
  const double4* s; double* out;
  FVM_DEV void operator()(long long i) const {
    const double4 v = s[i];
    out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
  }
};

// ================================================================= host side
System* systemCreate(Mesh* m) {
  requireReady();
  if (!m->hasGeometry) fail("system: mesh geometry not set (fvmgpu_mesh_set_geometry)");
  std::unique_ptr<System> s(new System);
  s->mesh = m;
  s->nSelf = m->nSelf;
  s->nTotal = m->nTotal;
  s->nnz = m->nnz;
  s->row = m->row.p;
  s->col = m->col.p;
  const size_t nt = m->nTotal;
  s->diag.alloc(nt); s->b.alloc(nt); s->delta.alloc(nt); s->x.alloc(nt);
  s->off.alloc(m->nnz);
  s->isBoundary.alloc(nt);
  s->diffusivity.alloc(nt); s->source.alloc(nt);
  s->cellState.alloc(nt);
  s->diag.zero(); s->b.zero(); s->delta.zero(); s->x.zero(); s->off.zero(); s->isBoundary.zero();
  s->source.zero();
  s->cellState.zero();
  parallelFor((long long)nt, FillKernel{s->diffusivity.p, 1.0});
  const size_t nb = m->nFaces - m->nInteriorFaces;
  s->bflux.alloc(nb + 1); s->rflux.alloc(nb + 1); s->coeffL.alloc(nb + 1); s->coeffR.alloc(nb + 1);
  s->bflux.zero(); s->rflux.zero(); s->coeffL.zero(); s->coeffR.zero();
  for (const FaceGroup& g : m->groups) {
    BcEntry e;
    e.offset = g.offset; e.count = g.count; e.kind = -1; e.groupKind = g.kind;
    e.p[0] = e.p[1] = e.p[2] = e.p[3] = 0.0;
    e.perFace = nullptr;
    if (g.kind == FVMGPU_GROUP_INTERFACE) e.kind = FVMGPU_BC_INTERFACE;  // applyInterfaceBC always runs
    s->bcs.push_back(e);
  }
  if (s->bcs.size() > MAX_BC_GROUPS) fail("system: more than %d face groups", MAX_BC_GROUPS);
  s->bcPerFace.resize(s->bcs.size());
  s->version = nextVersion();
  s->patternVersion = nextVersion();
  streamSync();
  return s.release();
}

System* systemCreateRaw(int nSelf, int nGhost, const int* row, const int* col, const double* diag,
                        const double* off, const double* b) {
  requireReady();
  std::unique_ptr<System> s(new System);
  const size_t nt = (size_t)nSelf + nGhost;
  s->nSelf = nSelf;
  s->nTotal = (int)nt;
  s->nnz = row[nt];
  s->rawRow.upload(row, nt + 1);
  s->rawCol.upload(col, s->nnz);
  s->row = s->rawRow.p;
  s->col = s->rawCol.p;
  s->diag.upload(diag, nt);
  s->off.upload(off, s->nnz);
  s->b.upload(b, nt);
  s->delta.alloc(nt); s->delta.zero();
  s->x.alloc(nt); s->x.zero();
  s->isBoundary.alloc(nt); s->isBoundary.zero();
  s->version = nextVersion();
  s->patternVersion = nextVersion();
  streamSync();
  return s.release();
}

static DBuf<double>* fieldBuf(System* s, int field, size_t& len) {
  const size_t nt = s->nTotal;
  const size_t nf = s->mesh ? s->mesh->nFaces : 0;
  switch (field) {
    case FVMGPU_FIELD_X: len = nt; return &s->x;
    case FVMGPU_FIELD_DIFFUSIVITY: len = nt; return &s->diffusivity;
    case FVMGPU_FIELD_SOURCE: len = nt; return &s->source;
    case FVMGPU_FIELD_FACE_FLUX: len = nf; return &s->faceFlux;
    case FVMGPU_FIELD_X_N1: len = nt; return &s->xN1;
    case FVMGPU_FIELD_X_N2: len = nt; return &s->xN2;
    case FVMGPU_FIELD_DENSITY: len = nt; return &s->density;
    case FVMGPU_FIELD_CONT_RESID: len = nt; return &s->contResid;
    case FVMGPU_FIELD_DELTA: len = nt; return &s->delta;
    case FVMGPU_FIELD_B: len = nt; return &s->b;
    default: return nullptr;
  }
}

void systemSetField(System* s, int field, const double* host, long long n, bool fill, double value) {
  requireReady();
  size_t len = 0;
  DBuf<double>* buf = fieldBuf(s, field, len);
  if (!buf) fail("set_field: field %d is not writable", field);
  if (!fill && (size_t)n != len) fail("set_field: field %d expects %zu values, got %lld", field, len, n);
  if (buf->n < len) buf->alloc(len);
  if (fill) parallelFor((long long)len, FillKernel{buf->p, value});
  else buf->upload(host, len);
  if (field == FVMGPU_FIELD_X) s->gradientValid = false;
  if (field == FVMGPU_FIELD_FACE_FLUX) s->hasFaceFlux = true;
  if (field == FVMGPU_FIELD_X_N1) s->hasXN1 = true;
  if (field == FVMGPU_FIELD_X_N2) s->hasXN2 = true;
  // b and delta do not enter the hierarchy or the ILU factors: the stamp stays (the reference keeps its
  // coarse levels while the matrix is unchanged, F/AMG.cpp:222-226)
}

void systemGetField(System* s, int field, double* host, long long n) {
  requireReady();
  if (field == FVMGPU_FIELD_GRADIENT) {
    if ((size_t)n != 3 * (size_t)s->nTotal) fail("get_field: gradient expects %zu values", 3 * (size_t)s->nTotal);
    DBuf<double> tmp(3 * (size_t)s->nTotal);
    parallelFor(s->nTotal, GradToAosKernel{s->cellState.p, tmp.p});
    tmp.download(host, tmp.n);
    return;
  }
  if (field == FVMGPU_FIELD_BFLUX_BOUNDARY) {
    if (!s->mesh) fail("get_field: raw systems have no boundary flux");
    const size_t nb = (size_t)(s->mesh->nFaces - s->mesh->nInteriorFaces);
    if ((size_t)n != nb) fail("get_field: boundary flux (boundary faces) expects %zu values", nb);
    s->bflux.download(host, nb);
    return;
  }
  if (field == FVMGPU_FIELD_BFLUX) {
    if (!s->mesh) fail("get_field: raw systems have no boundary flux");
    // laid out over ALL faces like the reference's per-group heatFlux arrays concatenated;
    // interior faces read 0
    const size_t nf = s->mesh->nFaces, ni = s->mesh->nInteriorFaces;
    if ((size_t)n != nf) fail("get_field: boundary flux expects %zu values", nf);
    std::memset(host, 0, ni * sizeof(double));
    s->bflux.download(host + ni, nf - ni);
    return;
  }
  size_t len = 0;
  DBuf<double>* buf = fieldBuf(s, field, len);
  if (!buf || !buf->p) fail("get_field: field %d not available", field);
  if ((size_t)n != len) fail("get_field: field %d has %zu values, asked for %lld", field, len, n);
  buf->download(host, len);
}

void systemSetBc(System* s, int groupId, int kind, const double* p, int np, const double* perFace) {
  requireReady();
  if (!s->mesh) fail("set_bc: raw systems have no boundary groups");
  for (size_t g = 0; g < s->bcs.size(); g++) {
    const FaceGroup& fg = s->mesh->groups[g];
    if (fg.id == groupId && fg.kind != FVMGPU_GROUP_INTERIOR) {
      BcEntry& e = s->bcs[g];
      e.kind = kind;
      for (int i = 0; i < 4; i++) e.p[i] = (i < np && p) ? p[i] : 0.0;
      if (perFace) {
        s->bcPerFace[g].upload(perFace, fg.count);
        e.perFace = s->bcPerFace[g].p;
      } else {
        e.perFace = nullptr;
      }
      s->bcsDirty = true;
      return;
    }
  }
  fail("set_bc: no boundary face group with id %d", groupId);
}

void systemHaloExchange(System* s, int field) {
  requireReady();
  if (!s->mesh) fail("halo_exchange: raw systems have no halo maps");
  size_t len = 0;
  DBuf<double>* buf = fieldBuf(s, field, len);
  if (!buf || !buf->p || len != (size_t)s->nTotal) fail("halo_exchange: field %d is not a cell field", field);
  s->mesh->halo.exchange(buf->p, 1);
  if (field == FVMGPU_FIELD_X) s->gradientValid = false;
}

void computeGradient(System* s) {
  requireReady();
  Mesh* m = s->mesh;
  if (!m) fail("compute_gradient: raw systems have no mesh");
  parallelFor(m->nTotal, GradientRows{m->nSelf, m->nInteriorFaces, m->row.p, m->col.p, m->entryFace.p,
                                      m->faceGroupOf.p, m->groupKindDev.p, m->faceGeom.p, s->x.p, m->gradW.p,
                                      m->nnz, s->cellState.p});
  // interface ghost cells: {gradient, x} of the owning rank (GradientModel::compute ends with the
  // gradient halo sync, F/GradientModel.h:600)
  m->halo.exchange(reinterpret_cast<double*>(s->cellState.p), 4);
  s->gradientValid = true;
}

void assemble(System* s, const fvmgpu_assemble_opts& o) {
  requireReady();
  Mesh* m = s->mesh;
  if (!m) fail("assemble: raw systems have no mesh");
  if (o.convection && !s->hasFaceFlux) fail("assemble: convection requested but FIELD_FACE_FLUX not set");
  if (o.time_order >= 1 && (!s->hasXN1 || !s->density.p)) fail("assemble: time derivative needs X_N1 and DENSITY");
  if (o.time_order >= 2 && !s->hasXN2) fail("assemble: second-order time derivative needs X_N2");
  if (o.time_order && !(o.dt > 0)) fail("assemble: dt must be positive");
  if (!s->gradientValid) computeGradient(s);
  if (s->bcsDirty) {
    s->bcsDev.upload(s->bcs.data(), s->bcs.size());
    s->bcsDirty = false;
  }
  AsmParams P;
  P.nSelf = m->nSelf; P.nTotal = m->nTotal; P.nInteriorFaces = m->nInteriorFaces;
  P.nGroups = (int)s->bcs.size();
  P.row = m->row.p; P.col = m->col.p; P.entryFace = m->entryFace.p; P.faceGroupOf = m->faceGroupOf.p;
  P.cellGeom = m->cellGeom.p; P.faceGeom = m->faceGeom.p; P.cellState = s->cellState.p;
  P.diffusivity = s->diffusivity.p; P.source = s->source.p;
  P.faceFlux = s->hasFaceFlux ? s->faceFlux.p : nullptr;
  P.xN1 = s->xN1.p; P.xN2 = s->xN2.p; P.density = s->density.p;
  P.contResid = s->contResid.p;
  P.bcs = s->bcsDev.p;
  P.x = s->x.p; P.diag = s->diag.p; P.off = s->off.p; P.b = s->b.p; P.isBoundary = s->isBoundary.p;
  P.bflux = s->bflux.p; P.rflux = s->rflux.p; P.coeffL = s->coeffL.p; P.coeffR = s->coeffR.p;
  P.o = o;
  parallelFor(m->nTotal, AssembleRows{P});
  // LinearSystem::initSolve: delta = 0
  s->delta.zero();
  s->version = nextVersion();
  if (o.apply_bcs) s->gradientValid = false;  // Dirichlet BCs rewrote x in the ghost cells
}

void postSolveUpdate(System* s) {
  requireReady();
  if (s->mesh) {
    Mesh* m = s->mesh;
    parallelFor(m->nTotal, PostSolveRows{m->nSelf, m->nInteriorFaces, m->row.p, m->col.p, m->entryFace.p,
                                         s->diag.p, s->off.p, s->b.p, s->isBoundary.p, s->delta.p, s->x.p,
                                         s->bflux.p, s->rflux.p, s->coeffL.p, s->coeffR.p, 1});
    m->halo.exchange(s->x.p, 1);  // updateSolution: x.sync() (F/LinearSystem.cpp:268)
  } else {
    parallelFor(s->nTotal, PostSolveRows{s->nSelf, 0, s->row, s->col, nullptr, s->diag.p, s->off.p, s->b.p,
                                         s->isBoundary.p, s->delta.p, s->x.p, nullptr, nullptr, nullptr, nullptr, 0});
  }
  s->gradientValid = false;
}

}  // namespace fvmgpu
