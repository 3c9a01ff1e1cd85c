/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product
 * path (fvm_b200/, libfvmgpu.so). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load the library built from this file.
 *
 * An independent, single-threaded, plain-C restatement of the reference's algorithm for the hot
 * path, written from the reference's behaviour, each function citing the file:line it follows
 * (F/ = src/fvm/src/modules/fvmbase/). It keeps the reference's SEQUENTIAL formulation: face-order
 * scatter assembly, lexicographic Gauss-Seidel, greedy row-order agglomeration -- i.e. it is NOT
 * the GPU algorithm. Parity status: PINNED -- tests/test_oracle_port.py checks it against the
 * reference's golden vectors (T/testLinearSolver.out levels + residuals, T/THERMAL_MATRIX/GOLDEN
 * matrix/rhs, T/AMG_MERGING_THERMAL/proc1 history) and, where oracle/_ref exists, bit-for-bit
 * against the reference compiled in place.
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, like the reference's x86-64 build).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ mesh view */
typedef struct {
  int dim, nSelf, nTotal, nFaces, nGroups;
  const int* faceCells;   /* 2F */
  const int* row;         /* cellCells, Nt+1 */
  const int* col;
  const int* pairToCol;   /* 2F (F/CRConnectivity.cpp:729-792) */
  const int* groupOffset; /* per group */
  const int* groupCount;
  const int* groupKind;   /* 0 interior 1 boundary 2 interface 3 symmetry */
  const double* faceArea; /* 3F */
  const double* faceAreaMag;
  const double* cellCentroid; /* 3Nt */
  const double* cellVolume;
} fvmo_mesh;

/* pair -> column map: position of (c0,c1) and (c1,c0) in cellCells */
void fvmo_pair_to_col(int nFaces, const int* faceCells, const int* row, const int* col, int* p2c) {
  for (int f = 0; f < nFaces; f++) {
    const int c0 = faceCells[2 * f], c1 = faceCells[2 * f + 1];
    p2c[2 * f] = p2c[2 * f + 1] = -1;
    for (int k = row[c0]; k < row[c0 + 1]; k++) if (col[k] == c1) { p2c[2 * f] = k; break; }
    for (int k = row[c1]; k < row[c1 + 1]; k++) if (col[k] == c0) { p2c[2 * f + 1] = k; break; }
  }
}

/* ------------------------------------------------------------------ gradient
 * GradientModel::getLeastSquaresGradientMatrix3D / 2D, F/GradientModel.h:126-282 / :284-436 */
void fvmo_ls_weights(const fvmo_mesh* m, double* coeffs /* 3*nnz AoS */) {
  const int nnz = m->row[m->nTotal];
  memset(coeffs, 0, sizeof(double) * 3 * (size_t)nnz);
  char* degenerate = (char*)calloc((size_t)m->nTotal, 1);
  for (int f = 0; f < m->nFaces; f++) {
    const int c0 = m->faceCells[2 * f], c1 = m->faceCells[2 * f + 1];
    double ds[3], mag = 0;
    for (int k = 0; k < 3; k++) ds[k] = m->cellCentroid[3 * c1 + k] - m->cellCentroid[3 * c0 + k];
    mag = sqrt(ds[0] * ds[0] + ds[1] * ds[1] + ds[2] * ds[2]);
    for (int k = 0; k < 3; k++) {
      coeffs[3 * m->pairToCol[2 * f] + k] = ds[k] / mag;
      coeffs[3 * m->pairToCol[2 * f + 1] + k] = (-ds[k]) / mag;
    }
  }
  for (int nc = 0; nc < m->nSelf; nc++) {
    double Ixx = 0, Iyy = 0, Izz = 0, Ixy = 0, Ixz = 0, Iyz = 0;
    for (int inb = m->row[nc]; inb < m->row[nc + 1]; inb++) {
      const double* ds = coeffs + 3 * inb;
      Ixx += ds[0] * ds[0]; Iyy += ds[1] * ds[1]; Ixy += ds[0] * ds[1];
      if (m->dim == 3) { Izz += ds[2] * ds[2]; Ixz += ds[0] * ds[2]; Iyz += ds[1] * ds[2]; }
    }
    if (m->dim == 3) {
      const double det = Ixx * (Iyy * Izz - Iyz * Iyz) - Ixy * (Ixy * Izz - Iyz * Ixz) + Ixz * (Ixy * Iyz - Iyy * Ixz);
      if (det > 1e-6) {
        const double Kxx = (Iyy * Izz - Iyz * Iyz) / det, Kxy = -(Ixy * Izz - Iyz * Ixz) / det;
        const double Kxz = (Ixy * Iyz - Iyy * Ixz) / det, Kyy = (Ixx * Izz - Ixz * Ixz) / det;
        const double Kyz = -(Ixx * Iyz - Ixy * Ixz) / det, Kzz = (Ixx * Iyy - Ixy * Ixy) / det;
        for (int inb = m->row[nc]; inb < m->row[nc + 1]; inb++) {
          double* c = coeffs + 3 * inb;
          const double d0 = c[0], d1 = c[1], d2 = c[2];
          c[0] = (Kxx * d0 + Kxy * d1 + Kxz * d2);
          c[1] = (Kxy * d0 + Kyy * d1 + Kyz * d2);
          c[2] = (Kxz * d0 + Kyz * d1 + Kzz * d2);
        }
      } else degenerate[nc] = 1;
    } else {
      const double det = Ixx * Iyy - Ixy * Ixy;
      if (det > 1e-26) {
        const double Kxx = Iyy / det, Kxy = -Ixy / det, Kyy = Ixx / det;
        for (int inb = m->row[nc]; inb < m->row[nc + 1]; inb++) {
          double* c = coeffs + 3 * inb;
          const double d0 = c[0], d1 = c[1];
          c[0] = (Kxx * d0 + Kxy * d1);
          c[1] = (Kxy * d0 + Kyy * d1);
          c[2] = 0;
        }
      } else degenerate[nc] = 1;
    }
  }
  for (int f = 0; f < m->nFaces; f++) {
    const int c0 = m->faceCells[2 * f], c1 = m->faceCells[2 * f + 1];
    double ds[3];
    for (int k = 0; k < 3; k++) ds[k] = m->cellCentroid[3 * c1 + k] - m->cellCentroid[3 * c0 + k];
    const double mag = sqrt(ds[0] * ds[0] + ds[1] * ds[1] + ds[2] * ds[2]);
    for (int k = 0; k < 3; k++) {
      coeffs[3 * m->pairToCol[2 * f] + k] /= mag;
      coeffs[3 * m->pairToCol[2 * f + 1] + k] /= mag;
    }
  }
  for (int f = 0; f < m->nFaces; f++) {
    const int c0 = m->faceCells[2 * f], c1 = m->faceCells[2 * f + 1];
    for (int k = 0; k < 3; k++) {
      if (degenerate[c0]) coeffs[3 * m->pairToCol[2 * f] + k] = 0.5 * m->faceArea[3 * f + k] / m->cellVolume[c0];
      if (degenerate[c1]) coeffs[3 * m->pairToCol[2 * f + 1] + k] = -0.5 * m->faceArea[3 * f + k] / m->cellVolume[c1];
    }
  }
  free(degenerate);
}

/* GradientMatrix::getGradient (F/GradientMatrix.h:55-76) + boundary copy / reflection
 * (F/GradientModel.h:530-566, reflectGradient :21-28) */
void fvmo_gradient(const fvmo_mesh* m, const double* coeffs, const double* x, double* grad /* 3Nt */) {
  for (int nr = 0; nr < m->nSelf; nr++) {
    double g[3] = {0, 0, 0};
    for (int nb = m->row[nr]; nb < m->row[nr + 1]; nb++) {
      const double v = x[m->col[nb]] - x[nr];
      for (int k = 0; k < 3; k++) g[k] += coeffs[3 * nb + k] * v;
    }
    for (int k = 0; k < 3; k++) grad[3 * nr + k] = g[k];
  }
  for (int gi = 0; gi < m->nGroups; gi++) {
    if (m->groupKind[gi] != 1 && m->groupKind[gi] != 3) continue;
    for (int f = m->groupOffset[gi]; f < m->groupOffset[gi] + m->groupCount[gi]; f++) {
      const int c0 = m->faceCells[2 * f], c1 = m->faceCells[2 * f + 1];
      if (m->groupKind[gi] == 3) {
        double en[3], dot = 0;
        for (int k = 0; k < 3; k++) en[k] = m->faceArea[3 * f + k] / m->faceAreaMag[f];
        for (int k = 0; k < 3; k++) dot += grad[3 * c0 + k] * en[k];
        const double t = 2.0 * dot;
        for (int k = 0; k < 3; k++) grad[3 * c1 + k] = grad[3 * c0 + k] - t * en[k];
      } else {
        for (int k = 0; k < 3; k++) grad[3 * c1 + k] = grad[3 * c0 + k];
      }
    }
  }
}

/* ------------------------------------------------------------------ assembly */
typedef struct {
  int kind;      /* -1 none, 0 dirichlet, 1 neumann, 2 extrapolation, 3 convective, 4 radiative, 5 mixed,
                    6 interface, 7 dirichlet-or-outflow */
  double p[4];
} fvmo_bc;

typedef struct {
  int diffusion, convection, source, time_order;
  double dt, underrelax;
  int apply_bcs, eliminate_boundary;
} fvmo_opts;

static double harmonic_average(double x0, double x1) { /* F/DiffusionDiscretization.h:19-27 */
  const double sum = x0 + x1;
  if (x0 + x1 != 0.0) return 2.0 * x0 * x1 / sum;
  return sum;
}

/* Linearizer order for ThermalModel (F/ThermalModel_impl.h:236-296): diffusion, convection,
 * source, [time derivative]; then the BC loop (:299-383); then LinearSystem::initSolve. */
void fvmo_thermal_assemble(const fvmo_mesh* m, const fvmo_opts* o, const fvmo_bc* bcs /* per group */,
                           double* x, const double* grad, const double* diffusivity, const double* source,
                           const double* faceFlux, const double* xN1, const double* xN2, const double* density,
                           const double* contResid, double* diag, double* off, double* r, int* isBoundary,
                           double* bflux, double* rflux, double* coeffL, double* coeffR /* all nFaces */) {
  const int nnz = m->row[m->nTotal];
  memset(diag, 0, sizeof(double) * (size_t)m->nTotal);
  memset(r, 0, sizeof(double) * (size_t)m->nTotal);
  memset(off, 0, sizeof(double) * (size_t)nnz);
  memset(isBoundary, 0, sizeof(int) * (size_t)m->nTotal);
  memset(bflux, 0, sizeof(double) * (size_t)m->nFaces);
  memset(rflux, 0, sizeof(double) * (size_t)m->nFaces);
  memset(coeffL, 0, sizeof(double) * (size_t)m->nFaces);
  memset(coeffR, 0, sizeof(double) * (size_t)m->nFaces);
  const int* p2c = m->pairToCol;
  if (o->diffusion) { /* F/DiffusionDiscretization.h:165-228 */
    for (int f = 0; f < m->nFaces; f++) {
      const int c0 = m->faceCells[2 * f], c1 = m->faceCells[2 * f + 1];
      const double vol0 = m->cellVolume[c0], vol1 = m->cellVolume[c1];
      double ds[3];
      for (int k = 0; k < 3; k++) ds[k] = m->cellCentroid[3 * c1 + k] - m->cellCentroid[3 * c0 + k];
      double fd;
      if (vol0 == 0.) fd = diffusivity[c1];
      else if (vol1 == 0.) fd = diffusivity[c0];
      else fd = harmonic_average(diffusivity[c0], diffusivity[c1]);
      const double* A = m->faceArea + 3 * f;
      const double am = m->faceAreaMag[f];
      const double diffMetric = am * am / (A[0] * ds[0] + A[1] * ds[1] + A[2] * ds[2]);
      const double diffCoeff = fd * diffMetric;
      double sec = 0.0;
      for (int k = 0; k < 3; k++) {
        const double sc = fd * (A[k] - ds[k] * diffMetric);
        const double gf = (grad[3 * c0 + k] * vol0 + grad[3 * c1 + k] * vol1) / (vol0 + vol1);
        sec += gf * sc;
      }
      const double dFlux = diffCoeff * (x[c1] - x[c0]) + sec;
      r[c0] += dFlux;
      r[c1] -= dFlux;
      off[p2c[2 * f]] += diffCoeff;
      off[p2c[2 * f + 1]] += diffCoeff;
      diag[c0] -= diffCoeff;
      diag[c1] -= diffCoeff;
    }
  }
  if (o->convection) { /* F/ConvectionDiscretization.h:119-199 */
    for (int f = 0; f < m->nFaces; f++) {
      const int c0 = m->faceCells[2 * f], c1 = m->faceCells[2 * f + 1];
      const double flux = faceFlux[f];
      double varFlux;
      if (o->convection == 2) varFlux = 0.5 * flux * (x[c0] + x[c0]); /* sic, :131 */
      else varFlux = (flux > 0.0) ? flux * x[c0] : flux * x[c1];
      if (flux > 0.0) { diag[c0] -= flux; off[p2c[2 * f + 1]] += flux; }
      else { diag[c1] += flux; off[p2c[2 * f]] -= flux; }
      r[c0] -= varFlux;
      r[c1] += varFlux;
    }
    if (contResid)
      for (int c = 0; c < m->nSelf; c++) diag[c] += contResid[c];
  }
  if (o->source && source) /* F/SourceDiscretization.h:54-57 */
    for (int c = 0; c < m->nSelf; c++) r[c] += m->cellVolume[c] * source[c];
  if (o->time_order == 1) { /* F/TimeDerivativeDiscretization.h:149-155 */
    for (int c = 0; c < m->nSelf; c++) {
      const double rhoVbydT = density[c] * m->cellVolume[c] / o->dt;
      r[c] -= rhoVbydT * (x[c] - xN1[c]);
      diag[c] -= rhoVbydT;
    }
  } else if (o->time_order == 2) { /* :102-108 */
    for (int c = 0; c < m->nSelf; c++) {
      const double rhoVbydT = density[c] * m->cellVolume[c] / o->dt;
      r[c] -= rhoVbydT * (1.5 * x[c] - 2.0 * xN1[c] + 0.5 * xN2[c]);
      diag[c] -= rhoVbydT * 1.5;
    }
  }
  if (o->apply_bcs) { /* GenericBCS, F/GenericBCS.h:77-356 */
    const double sb = 5.670373E-8;
    for (int gi = 0; gi < m->nGroups; gi++) {
      if (bcs[gi].kind < 0) continue;
      for (int f = m->groupOffset[gi]; f < m->groupOffset[gi] + m->groupCount[gi]; f++) {
        const int c0 = m->faceCells[2 * f], c1 = m->faceCells[2 * f + 1];
        const double am = m->faceAreaMag[f];
        const double* p = bcs[gi].p;
        int kind = bcs[gi].kind;
        if (kind == 7) kind = (faceFlux && faceFlux[f] > 0.) ? 2 : 0; /* F/ThermalModel_impl.h:313-331 */
        double* c01 = off + p2c[2 * f];
        double* c10 = off + p2c[2 * f + 1];
        if (kind == 0) { /* applyDirichletBC :77-115 */
          const double fluxB = -r[c1], dFluxdXC0 = -*c10, dFluxdXC1 = -diag[c1], dRC0dXC1 = *c01;
          const double dXC1 = p[0] - x[c1];
          const double dFlux = dFluxdXC1 * dXC1, dRC0 = dRC0dXC1 * dXC1;
          r[c0] += dRC0;
          *c01 = 0.0;
          x[c1] = p[0];
          *c10 = 0.0;
          r[c1] = 0.0;
          diag[c1] = -1.0;
          coeffL[f] = dFluxdXC0; coeffR[f] = 0.0; bflux[f] = fluxB; rflux[f] = dFlux;
        } else if (kind == 1) { /* applyNeumannBC :129-157 */
          const double fluxB = -r[c1];
          r[c1] = p[0] * am - fluxB;
          isBoundary[c1] = 1;
          bflux[f] = p[0] * am;
        } else if (kind == 2) { /* applyExtrapolationBC :180-212 */
          const double fluxB = -r[c1], dFluxdXC0 = -*c10, dFluxdXC1 = -diag[c1];
          const double xc0mxc1 = x[c0] - x[c1];
          diag[c0] += dFluxdXC1;
          r[c0] += dFluxdXC1 * xc0mxc1;
          *c01 = 0;
          diag[c1] = -1.0;
          *c10 = 1.0;
          r[c1] = xc0mxc1;
          isBoundary[c1] = 1;
          coeffL[f] = dFluxdXC0; coeffR[f] = dFluxdXC0; bflux[f] = fluxB; rflux[f] = 0;
        } else if (kind == 3) { /* applyConvectionBC :214-245 */
          const double fluxInterior = -r[c1];
          const double fluxBoundary = -p[0] * (x[c1] - p[1]) * am;
          r[c1] = fluxBoundary - fluxInterior;
          diag[c1] -= p[0] * am;
          isBoundary[c1] = 1;
          bflux[f] = fluxBoundary; rflux[f] = 0; coeffL[f] = 0; coeffR[f] = -p[0] * am;
        } else if (kind == 4) { /* applyRadiationBC :253-288 */
          const double xb = x[c1], Xinf = p[1];
          const double fluxInterior = -r[c1];
          const double fluxBoundary = -p[0] * sb * (xb * xb * xb * xb - Xinf * Xinf * Xinf * Xinf) * am;
          r[c1] = fluxBoundary - fluxInterior;
          diag[c1] -= 4 * p[0] * sb * xb * xb * xb * am;
          isBoundary[c1] = 1;
          bflux[f] = fluxBoundary; rflux[f] = 0; coeffL[f] = 0; coeffR[f] = -4 * p[0] * sb * xb * xb * xb * am;
        } else if (kind == 5) { /* applyMixedBC :290-323 */
          const double xb = x[c1], h = p[0], em = p[1], Xinf = p[2];
          const double fluxInterior = -r[c1];
          const double fluxBoundary = (-em * sb * (xb * xb * xb * xb - Xinf * Xinf * Xinf * Xinf) - h * (xb - Xinf)) * am;
          r[c1] = fluxBoundary - fluxInterior;
          diag[c1] -= (4 * em * sb * xb * xb * xb + h) * am;
          isBoundary[c1] = 1;
          bflux[f] = fluxBoundary; rflux[f] = 0; coeffL[f] = 0; coeffR[f] = -4 * em * sb * xb * xb * xb * am;
        } else if (kind == 6) { /* applyInterfaceBC :325-356, ghost is c1 */
          const double fluxInterior = -r[c1];
          coeffL[f] = -1.0 * *c10; coeffR[f] = 1.0 * *c01;
          r[c1] = 0;
          *c10 = 0.0;
          bflux[f] = fluxInterior; rflux[f] = 0;
        }
      }
    }
  }
  if (o->underrelax > 0) /* F/Underrelaxer.h:49-52 */
    for (int c = 0; c < m->nSelf; c++) diag[c] /= o->underrelax;
  if (o->eliminate_boundary) { /* CRMatrix::eliminateBoundaryEquations -> eliminateRow, F/CRMatrix.h:899-944,1064-1085 */
    for (int j = m->nSelf; j < m->nTotal; j++) {
      if (!isBoundary[j]) continue;
      const double a_jj = diag[j];
      for (int nb = m->row[j]; nb < m->row[j + 1]; nb++) {
        const int i = m->col[nb];
        double* a_ij = NULL;
        for (int k = m->row[i]; k < m->row[i + 1]; k++) if (m->col[k] == j) { a_ij = off + k; break; }
        for (int nb2 = m->row[j]; nb2 < m->row[j + 1]; nb2++) {
          const int k = m->col[nb2];
          const double a_jk = off[nb2];
          if (i != k) {
            for (int q = m->row[i]; q < m->row[i + 1]; q++)
              if (m->col[q] == k) { off[q] -= *a_ij * (a_jk / a_jj); break; }
          } else {
            diag[i] -= *a_ij * (a_jk / a_jj);
          }
        }
        r[i] -= *a_ij * (r[j] / a_jj);
        *a_ij = 0.0;
      }
    }
  }
}

/* LinearSystem::postSolve + updateSolution (F/LinearSystem.cpp:250-269): CRMatrix::solveBoundary
 * (F/CRMatrix.h:433-454), then the boundary-flux rows (FluxJacobianMatrix / DiagonalMatrix = -1) */
void fvmo_post_solve(const fvmo_mesh* m, const double* diag, const double* off, const double* b,
                     const int* isBoundary, double* delta, double* x, double* bflux, const double* rflux,
                     const double* coeffL, const double* coeffR) {
  for (int nr = m->nSelf; nr < m->nTotal; nr++)
    if (isBoundary[nr]) {
      double sum = b[nr];
      for (int nb = m->row[nr]; nb < m->row[nr + 1]; nb++) sum += off[nb] * delta[m->col[nb]];
      delta[nr] = -sum / diag[nr];
    }
  for (int gi = 0; gi < m->nGroups; gi++) {
    if (m->groupKind[gi] == 0) continue;
    for (int f = m->groupOffset[gi]; f < m->groupOffset[gi] + m->groupCount[gi]; f++) {
      const int c0 = m->faceCells[2 * f], c1 = m->faceCells[2 * f + 1];
      double rr = rflux[f];
      rr += coeffL[f] * delta[c0] + coeffR[f] * delta[c1];
      bflux[f] += -rr / -1.0;
    }
  }
  for (int c = 0; c < m->nTotal; c++) x[c] += delta[c];
}

/* ------------------------------------------------------------------ AMG (F/AMG.cpp, F/CRMatrix.h) */
typedef struct {
  int nSelf, nTotal;
  int *row, *col;
  double *diag, *off, *b, *x, *r;
  int* isBoundary; /* may be NULL */
  int* ci;         /* coarse index per row (nTotal), -1 = none */
  int owns;
} fvmo_level;

typedef struct {
  int nMaxIterations, verbosity;
  double relativeTolerance, absoluteTolerance;
  int maxCoarseLevels, nPreSweeps, nPostSweeps, coarseGroupSize;
  double weightRatioThreshold;
  int cycleType, smootherType;
} fvmo_amg_opts;

/* CRMatrix::createCoarsening, F/CRMatrix.h:468-586 */
static int create_coarsening(const fvmo_level* L, int groupSize, double thr, int* coarseIndex) {
  const int nRows = L->nSelf;
  for (int i = 0; i < L->nTotal; i++) coarseIndex[i] = -1;
  int nCoarseRows = 0;
  int* coarseCount = (int*)calloc((size_t)(nRows > 0 ? nRows : 1), sizeof(int));
  for (int nr = 0; nr < nRows; nr++) {
    if (!(coarseIndex[nr] == -1 && !(L->isBoundary && L->isBoundary[nr]))) continue;
    int current = nr, colMaxGrouped = -1, colMaxUngrouped = -1, nGrouped;
    coarseIndex[current] = nCoarseRows;
    for (nGrouped = 1; nGrouped < groupSize; nGrouped++) {
      double maxWeightUngrouped = 0, maxWeightGrouped = 0;
      colMaxGrouped = -1;
      colMaxUngrouped = -1;
      for (int nb = L->row[current]; nb < L->row[current + 1]; nb++) {
        const int nc = L->col[nb];
        if (nc < nRows && !(L->isBoundary && L->isBoundary[nc])) {
          const double d0 = fabs(L->diag[nr]), d1 = fabs(L->diag[nc]); /* doubleMeasure = fabs, F/NumType.h:106 */
          const double thisWeight = fabs(fabs(L->off[nb]) / (d0 > d1 ? d0 : d1));
          if (coarseIndex[nc] == -1) {
            if (colMaxUngrouped == -1 || (thisWeight > maxWeightUngrouped)) { colMaxUngrouped = nc; maxWeightUngrouped = thisWeight; }
          } else if (coarseIndex[nc] != coarseIndex[nr]) {
            if (colMaxGrouped == -1 || (thisWeight > maxWeightGrouped)) { colMaxGrouped = nc; maxWeightGrouped = thisWeight; }
          }
        }
      }
      if ((colMaxUngrouped != -1) && (colMaxGrouped == -1 || (maxWeightUngrouped > thr * maxWeightGrouped))) {
        coarseIndex[colMaxUngrouped] = coarseIndex[current];
        coarseCount[coarseIndex[current]]++;
        current = colMaxUngrouped;
      } else break;
    }
    if (nGrouped > 1 || colMaxGrouped == -1 || coarseCount[coarseIndex[colMaxGrouped]] > groupSize + 2) {
      coarseCount[coarseIndex[nr]]++;
      nCoarseRows++;
    } else {
      coarseIndex[nr] = coarseIndex[colMaxGrouped];
      coarseCount[coarseIndex[colMaxGrouped]]++;
    }
  }
  free(coarseCount);
  return nCoarseRows;
}

/* createCoarseToFineMapping (F/MultiFieldMatrix.cpp:626-655), createCoarseConnectivity
 * (F/CRMatrix.h:597-691), createCoarseMatrix (:699-758) */
static fvmo_level* create_coarse(const fvmo_level* F, const int* ci, int nc) {
  fvmo_level* C = (fvmo_level*)calloc(1, sizeof(fvmo_level));
  C->nSelf = C->nTotal = nc;
  C->owns = 1;
  int* c2fRow = (int*)calloc((size_t)nc + 1, sizeof(int));
  for (int nr = 0; nr < F->nTotal; nr++) if (ci[nr] >= 0) c2fRow[ci[nr] + 1]++;
  for (int i = 0; i < nc; i++) c2fRow[i + 1] += c2fRow[i];
  int* c2f = (int*)malloc(sizeof(int) * (size_t)(c2fRow[nc] > 0 ? c2fRow[nc] : 1));
  int* fill = (int*)calloc((size_t)nc + 1, sizeof(int));
  for (int nr = 0; nr < F->nTotal; nr++) if (ci[nr] >= 0) c2f[c2fRow[ci[nr]] + fill[ci[nr]]++] = nr;
  char* counted = (char*)calloc((size_t)nc + 1, 1);
  C->row = (int*)calloc((size_t)nc + 1, sizeof(int));
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) {
      for (int i = 0; i < nc; i++) C->row[i + 1] += C->row[i];
      C->col = (int*)malloc(sizeof(int) * (size_t)(C->row[nc] > 0 ? C->row[nc] : 1));
      memset(fill, 0, sizeof(int) * ((size_t)nc + 1));
    }
    for (int I = 0; I < nc; I++) {
      for (int q = c2fRow[I]; q < c2fRow[I + 1]; q++) {
        const int nrFine = c2f[q];
        for (int nb = F->row[nrFine]; nb < F->row[nrFine + 1]; nb++) {
          const int J = ci[F->col[nb]];
          if (J >= 0 && I != J && !counted[J]) {
            counted[J] = 1;
            if (pass == 0) C->row[I + 1]++;
            else C->col[C->row[I] + fill[I]++] = J;
          }
        }
      }
      for (int q = c2fRow[I]; q < c2fRow[I + 1]; q++) {
        const int nrFine = c2f[q];
        for (int nb = F->row[nrFine]; nb < F->row[nrFine + 1]; nb++) {
          const int J = ci[F->col[nb]];
          if (J >= 0) counted[J] = 0;
        }
      }
    }
  }
  const int cnnz = C->row[nc];
  C->diag = (double*)calloc((size_t)nc + 1, sizeof(double));
  C->off = (double*)calloc((size_t)(cnnz > 0 ? cnnz : 1), sizeof(double));
  int* pos = (int*)calloc((size_t)nc + 1, sizeof(int));
  for (int I = 0; I < nc; I++) {
    for (int nb = C->row[I]; nb < C->row[I + 1]; nb++) pos[C->col[nb]] = nb;
    for (int q = c2fRow[I]; q < c2fRow[I + 1]; q++) {
      const int nrFine = c2f[q];
      C->diag[I] += F->diag[nrFine];
      for (int nb = F->row[nrFine]; nb < F->row[nrFine + 1]; nb++) {
        const int J = ci[F->col[nb]];
        if (J < 0) continue;
        if (I != J) C->off[pos[J]] += F->off[nb];
        else C->diag[I] += F->off[nb];
      }
    }
  }
  C->b = (double*)calloc((size_t)nc + 1, sizeof(double));
  C->x = (double*)calloc((size_t)nc + 1, sizeof(double));
  C->r = (double*)calloc((size_t)nc + 1, sizeof(double));
  free(pos); free(counted); free(fill); free(c2f); free(c2fRow);
  return C;
}

static void free_level(fvmo_level* L) {
  if (!L) return;
  if (L->owns) { free(L->row); free(L->col); free(L->diag); free(L->off); free(L->b); free(L->x); free(L->r); }
  free(L->ci);
  free(L);
}

static void forward_gs(const fvmo_level* L) { /* F/CRMatrix.h:303-322 */
  for (int nr = 0; nr < L->nSelf; nr++) {
    double sum = L->b[nr];
    for (int nb = L->row[nr]; nb < L->row[nr + 1]; nb++) sum += L->off[nb] * L->x[L->col[nb]];
    L->x[nr] = -sum / L->diag[nr];
  }
}
static void reverse_gs(const fvmo_level* L) { /* :329-346 */
  for (int nr = L->nSelf - 1; nr >= 0; nr--) {
    double sum = L->b[nr];
    for (int nb = L->row[nr]; nb < L->row[nr + 1]; nb++) sum += L->off[nb] * L->x[L->col[nb]];
    L->x[nr] = -sum / L->diag[nr];
  }
}
static void jacobi(const fvmo_level* L) { /* :353-374 via MultiFieldMatrix::Jacobi (new values through a temp) */
  for (int nr = 0; nr < L->nSelf; nr++) {
    double sum = L->b[nr];
    for (int nb = L->row[nr]; nb < L->row[nr + 1]; nb++) sum += L->off[nb] * L->x[L->col[nb]];
    L->r[nr] = -sum / L->diag[nr];
  }
  for (int nr = 0; nr < L->nSelf; nr++) L->x[nr] = L->r[nr];
}
static void compute_residual(const fvmo_level* L) { /* :407-426 */
  for (int nr = 0; nr < L->nSelf; nr++) {
    double v = L->b[nr] + L->diag[nr] * L->x[nr];
    for (int nb = L->row[nr]; nb < L->row[nr + 1]; nb++) v += L->off[nb] * L->x[L->col[nb]];
    L->r[nr] = v;
  }
}
static double one_norm(const double* a, int n) { /* F/Array.h:290-299 */
  double s = 0;
  for (int i = 0; i < n; i++) s += fabs(a[i]);
  return s;
}

typedef struct {
  fvmo_level* lv[64];
  int n;
  fvmo_amg_opts o;
} fvmo_hier;

static void do_sweeps(fvmo_hier* H, int nSweeps, int l) { /* F/AMG.cpp:43-68 */
  for (int i = 0; i < nSweeps; i++) {
    if (H->o.smootherType == 0) { forward_gs(H->lv[l]); reverse_gs(H->lv[l]); }
    else { jacobi(H->lv[l]); jacobi(H->lv[l]); }
  }
}
static void cycle(fvmo_hier* H, int type, int l) { /* F/AMG.cpp:70-147 */
  do_sweeps(H, H->o.nPreSweeps, l);
  if (l + 1 < H->n) {
    fvmo_level *F = H->lv[l], *C = H->lv[l + 1];
    compute_residual(F);
    memset(C->b, 0, sizeof(double) * (size_t)C->nTotal);
    memset(C->x, 0, sizeof(double) * (size_t)C->nTotal);
    for (int i = 0; i < F->nSelf; i++) if (F->ci[i] >= 0) C->b[F->ci[i]] += F->r[i]; /* Array::inject, F/Array.h:427-436 */
    cycle(H, type, l + 1);
    if (type == 1) cycle(H, 1, l + 1);
    else if (type == 2) cycle(H, 0, l + 1);
    for (int i = 0; i < F->nSelf; i++) if (F->ci[i] >= 0) F->x[i] += C->x[F->ci[i]]; /* Array::correct :438-467 */
  }
  do_sweeps(H, H->o.nPostSweeps, l);
}

/* AMG::createCoarseLevels, parallel-build order (F/AMG.cpp:149-210): push, then stop at <= 3 rows */
static void create_levels(fvmo_hier* H) {
  for (int n = 0; n < H->o.maxCoarseLevels && H->n < 63; n++) {
    fvmo_level* F = H->lv[H->n - 1];
    int* ci = (int*)malloc(sizeof(int) * (size_t)(F->nTotal > 0 ? F->nTotal : 1));
    const int nc = create_coarsening(F, H->o.coarseGroupSize, H->o.weightRatioThreshold, ci);
    if (nc == F->nTotal) { free(ci); break; } /* getLocalSize unchanged */
    fvmo_level* C = create_coarse(F, ci, nc);
    F->ci = ci;
    H->lv[H->n++] = C;
    if (nc <= 3) break;
  }
}

/* AMG::solve (F/AMG.cpp:219-282) or BCGStab::solve (F/BCGStab.cpp:26-170) on a CSR system with
 * separate diagonal. x holds delta on entry and exit. history[0..*nHist-1] = L1 residuals.
 * levelSizes: coarse level sizes, -1 terminated (cap 64). returns cycles / iterations. */
static void precondition(fvmo_hier* H, const double* rhs, double* out) {
  fvmo_level* L0 = H->lv[0];
  memcpy(L0->b, rhs, sizeof(double) * (size_t)L0->nSelf);
  memset(L0->x, 0, sizeof(double) * (size_t)L0->nTotal);
  cycle(H, H->o.cycleType, 0);
  memcpy(out, L0->x, sizeof(double) * (size_t)L0->nSelf);
}

/* CRMatrix::createCoarsening alone (F/CRMatrix.h:468-586) on a square CSR without ghost rows; signature of
 * fvmgpu_aggregate_fn (include/fvmgpu.h) so that the tests can hand it to the library's verification hook */
int fvmo_create_coarsening(void* user, int nRows, const int* row, const int* col, const double* diag,
                           const double* off, const int* isBoundary, int groupSize, double thr, int* coarseIndex) {
  (void)user;
  fvmo_level L;
  memset(&L, 0, sizeof(L));
  L.nSelf = nRows; L.nTotal = nRows;
  L.row = (int*)row; L.col = (int*)col; L.diag = (double*)diag; L.off = (double*)off;
  L.isBoundary = (int*)isBoundary;
  return create_coarsening(&L, groupSize, thr, coarseIndex);
}

int fvmo_solve(int nSelf, int nGhost, const int* row, const int* col, const double* diag, const double* off,
               const double* b, const int* isBoundary, const fvmo_amg_opts* o, int useBcgstab, double* x,
               double* history, int histCap, int* nHist, int* levelSizes) {
  fvmo_hier H;
  memset(&H, 0, sizeof(H));
  H.o = *o;
  const int nt = nSelf + nGhost;
  fvmo_level* L0 = (fvmo_level*)calloc(1, sizeof(fvmo_level));
  L0->nSelf = nSelf; L0->nTotal = nt;
  L0->row = (int*)row; L0->col = (int*)col; L0->diag = (double*)diag; L0->off = (double*)off;
  L0->isBoundary = (int*)isBoundary;
  L0->b = (double*)malloc(sizeof(double) * (size_t)nt);
  L0->x = x;
  L0->r = (double*)calloc((size_t)nt, sizeof(double));
  memcpy(L0->b, b, sizeof(double) * (size_t)nt);
  H.lv[0] = L0;
  H.n = 1;
  if (useBcgstab) { H.o.nMaxIterations = o->nMaxIterations; }
  create_levels(&H);
  if (levelSizes) { int k = 0; for (int l = 1; l < H.n; l++) levelSizes[k++] = H.lv[l]->nTotal; levelSizes[k] = -1; }
  int nh = 0, iters = 0;
  if (!useBcgstab) {
    compute_residual(L0);
    const double rNorm0 = one_norm(L0->r, nSelf);
    if (nh < histCap) history[nh++] = rNorm0;
    if (!(rNorm0 < o->absoluteTolerance)) {
      for (int i = 1; i < o->nMaxIterations; i++) {
        cycle(&H, o->cycleType, 0);
        iters++;
        compute_residual(L0);
        const double rNorm = one_norm(L0->r, nSelf);
        if (nh < histCap) history[nh++] = rNorm;
        if (rNorm < o->absoluteTolerance || rNorm / rNorm0 < o->relativeTolerance) break;
      }
    }
  } else {
    const size_t nb = sizeof(double) * (size_t)nt;
    double *xs = (double*)malloc(nb), *bOrig = (double*)malloc(nb), *r = (double*)calloc((size_t)nt, 8),
           *rT = (double*)calloc((size_t)nt, 8), *p = (double*)calloc((size_t)nt, 8), *pHat = (double*)calloc((size_t)nt, 8),
           *v = (double*)calloc((size_t)nt, 8), *t = (double*)calloc((size_t)nt, 8);
    memcpy(xs, x, nb);
    memcpy(bOrig, b, nb);
    compute_residual(L0);
    const double rNorm0 = one_norm(L0->r, nSelf);
    if (nh < histCap) history[nh++] = rNorm0;
    memcpy(r, L0->r, nb);
    memcpy(rT, L0->r, nb);
    double rho = 0, rhoPrev = 0, alpha = 0, omega = 0;
    int haveP = 0;
    for (int i = 0; i < o->nMaxIterations; i++) {
      iters++;
      rhoPrev = rho;
      rho = 0;
      for (int k = 0; k < nSelf; k++) rho += r[k] * rT[k];
      if (!haveP) { memcpy(p, r, nb); haveP = 1; }
      else {
        const double beta = (rho / rhoPrev) * (alpha / omega);
        for (int k = 0; k < nSelf; k++) { p[k] -= omega * v[k]; p[k] *= beta; p[k] += r[k]; }
      }
      precondition(&H, p, pHat);
      for (int nr = 0; nr < nSelf; nr++) { /* CRMatrix::multiply */
        double y = diag[nr] * pHat[nr];
        for (int q = row[nr]; q < row[nr + 1]; q++) y += off[q] * pHat[col[q]];
        v[nr] = y;
      }
      double rtv = 0;
      for (int k = 0; k < nSelf; k++) rtv += rT[k] * v[k];
      alpha = rho / rtv;
      for (int k = 0; k < nSelf; k++) { xs[k] -= alpha * pHat[k]; r[k] -= alpha * v[k]; }
      double rNorm = one_norm(r, nSelf);
      if (rNorm < o->absoluteTolerance) break;
      precondition(&H, r, pHat);
      for (int nr = 0; nr < nSelf; nr++) {
        double y = diag[nr] * pHat[nr];
        for (int q = row[nr]; q < row[nr + 1]; q++) y += off[q] * pHat[col[q]];
        t[nr] = y;
      }
      double tdotr = 0, tdott = 0;
      for (int k = 0; k < nSelf; k++) { tdotr += t[k] * r[k]; tdott += t[k] * t[k]; }
      omega = tdotr / tdott;
      for (int k = 0; k < nSelf; k++) { xs[k] -= omega * pHat[k]; r[k] -= omega * t[k]; }
      rNorm = one_norm(r, nSelf);
      if (nh < histCap) history[nh++] = rNorm;
      if (rNorm < o->absoluteTolerance || rNorm / rNorm0 < o->relativeTolerance) break;
    }
    memcpy(x, xs, nb);
    free(xs); free(bOrig); free(r); free(rT); free(p); free(pHat); free(v); free(t);
  }
  if (nHist) *nHist = nh;
  free(L0->b); free(L0->r);
  L0->owns = 0;
  for (int l = 0; l < H.n; l++) free_level(H.lv[l]);
  return iters;
}
