// fvm_b200 / libfvmgpu -- ILU(0) of the reference (CRMatrix::compute_ILU0 / lowerSolve / upperSolve,
// F/CRMatrix.h:1546-1715) and its two users: ILU0Solver::solve (F/ILU0Solver.cpp:46-93) and
// ILU0Solver::smooth as the preconditioner of BCGStab (F/BCGStab.cpp, T/PARALLEL_CAVITY_ILU0).
//
// The reference factorises and substitutes row after row. Row k only needs the rows of its LOWER
// entries (j < k), so the rows are grouped into dependency levels (level(k) = 1 + max level of its lower
// neighbours; for the upper solve the same with the higher neighbours) and one launch handles one level
// with one thread per row. Inside a row every operation happens in the reference's order -- lower
// entries in their stored (face) order, then the inverted diagonal -- so the factors and the solves are
// bit-identical to the sequential code (compiled with -fmad=false like the assembly). A natural-order hex
// mesh has nx + ny + nz - 2 levels: this is a latency-bound smoother by construction, kept for parity of
// the solver family, not for speed. The factorisation ignores ghost columns (j >= nSelf), as the
// reference does; across ranks it is therefore block-Jacobi ILU, again as in the reference.
#include "solver.cuh"

#include <algorithm>

namespace fvmgpu {

struct Ilu0 {
  int n = 0;
  long long nnz = 0;
  DBuf<int> row, col, diagIdx, src;   // ILU pattern (diagonal explicit); src: >= 0 offdiag index, < 0: ~row (diagonal)
  DBuf<double> coef, y;
  DBuf<int> orderL, orderU;           // rows sorted by level
  std::vector<int> startL, startU;    // level boundaries in orderL / orderU
  unsigned long long patternVersion = 0;   // System::patternVersion the pattern analysis was done for
  unsigned long long factoredVersion = 0;  // System::version the factors belong to (0: none)
};
void Ilu0Deleter::operator()(Ilu0* p) const { delete p; }

struct IluFillKernel {
  const int* src; const double* diag; const double* off; double* coef;
  FVM_DEV void operator()(long long p) const { const int s = src[p]; coef[p] = s >= 0 ? off[s] : diag[~s]; }
};
// one row of the IKJ factorisation, F/CRMatrix.h:1630-1668 (iw[] replaced by a search in the short row)
struct IluFactorRows {
  int first; const int* order; const int* row; const int* col; const int* diagIdx; double* coef;
  FVM_DEV void operator()(long long t) const {
    const int k = order[first + t];
    const int j1 = row[k], j2 = row[k + 1], jd = diagIdx[k];
    for (int j = j1; j < jd; j++) {
      const int jrow = col[j];
      const double t1 = coef[j] * coef[diagIdx[jrow]];
      coef[j] = t1;
      for (int jj = diagIdx[jrow] + 1; jj < row[jrow + 1]; jj++) {
        const int c = col[jj];
        for (int jw = j1; jw < j2; jw++)
          if (col[jw] == c) { coef[jw] -= t1 * coef[jj]; break; }
      }
    }
    coef[jd] = 1.0 / coef[jd];
  }
};
struct IluLowerRows {  // y_j = -b_j - sum_{k < diag} c_k y_col(k)
  int first; const int* order; const int* row; const int* col; const int* diagIdx; const double* coef; const double* b;
  double* y;
  FVM_DEV void operator()(long long t) const {
    const int j = order[first + t];
    double yj = -b[j];
    for (int k = row[j]; k < diagIdx[j]; k++) yj -= coef[k] * y[col[k]];
    y[j] = yj;
  }
};
struct IluUpperRows {  // x_j = c_diag * (y_j - sum_{k > diag} c_k x_col(k))
  int first; const int* order; const int* row; const int* col; const int* diagIdx; const double* coef; const double* y;
  double* x;
  FVM_DEV void operator()(long long t) const {
    const int j = order[first + t];
    double xj = y[j];
    for (int k = diagIdx[j] + 1; k < row[j + 1]; k++) xj -= coef[k] * x[col[k]];
    x[j] = coef[diagIdx[j]] * xj;
  }
};
struct CsrResidualRows {  // r = b + A x on the system's own CSR (natural numbering, ghost columns included)
  const int* row; const int* col; const double* diag; const double* off; const double* b; const double* x;
  FVM_DEV void operator()(long long i, double* o) const {
    double v = b[i] + diag[i] * x[i];
    for (int k = row[i]; k < row[i + 1]; k++) v += off[k] * x[col[k]];
    o[0] = fabs(v);
  }
};

// pattern + levels (host, once per connectivity), F/CRMatrix.h:1548-1612
static void iluPattern(Ilu0& I, System* sys) {
  const int n = sys->nSelf;
  std::vector<int> row((size_t)sys->nTotal + 1), col((size_t)sys->nnz);
  copyD2H(row.data(), sys->row, row.size() * sizeof(int));
  if (sys->nnz) copyD2H(col.data(), sys->col, col.size() * sizeof(int));
  std::vector<int> irow((size_t)n + 1, 0), icol, isrc, idiag((size_t)n);
  for (int r = 0; r < n; r++) {
    for (int nb = row[r]; nb < row[r + 1]; nb++)       // lower coefficients first, in stored order
      if (col[nb] < n && col[nb] < r) { icol.push_back(col[nb]); isrc.push_back(nb); }
    idiag[r] = (int)icol.size();
    icol.push_back(r); isrc.push_back(~r);
    for (int nb = row[r]; nb < row[r + 1]; nb++)       // then the upper ones
      if (col[nb] < n && col[nb] > r) { icol.push_back(col[nb]); isrc.push_back(nb); }
    irow[(size_t)r + 1] = (int)icol.size();
  }
  auto levelOrder = [&](bool lower, DBuf<int>& order, std::vector<int>& start) {
    std::vector<int> lvl((size_t)n, 0);
    int maxLvl = 0;
    if (lower) {
      for (int r = 0; r < n; r++) {
        int l = 0;
        for (int k = irow[r]; k < idiag[r]; k++) l = std::max(l, lvl[icol[k]] + 1);
        lvl[r] = l; maxLvl = std::max(maxLvl, l);
      }
    } else {
      for (int r = n - 1; r >= 0; r--) {
        int l = 0;
        for (int k = idiag[r] + 1; k < irow[(size_t)r + 1]; k++) l = std::max(l, lvl[icol[k]] + 1);
        lvl[r] = l; maxLvl = std::max(maxLvl, l);
      }
    }
    start.assign((size_t)maxLvl + 2, 0);
    for (int r = 0; r < n; r++) start[(size_t)lvl[r] + 1]++;
    for (size_t l = 1; l < start.size(); l++) start[l] += start[l - 1];
    std::vector<int> pos(start.begin(), start.end() - 1), ord((size_t)n);
    for (int r = 0; r < n; r++) ord[(size_t)pos[lvl[r]]++] = r;
    order.upload(ord.data(), ord.size());
  };
  I.n = n;
  I.nnz = (long long)icol.size();
  I.row.upload(irow.data(), irow.size());
  I.col.upload(icol.data(), icol.size());
  I.src.upload(isrc.data(), isrc.size());
  I.diagIdx.upload(idiag.data(), idiag.size());
  I.coef.alloc(icol.size());
  I.y.alloc((size_t)n);
  if (n) {
    levelOrder(true, I.orderL, I.startL);
    levelOrder(false, I.orderU, I.startU);
  }
  I.patternVersion = sys->patternVersion;
  I.factoredVersion = 0;
}

static void iluEnsure(Ilu0& I, System* sys) {
  // stamps, not addresses: a destroyed system's address (host and device) is readily handed out again
  if (I.patternVersion != sys->patternVersion || I.n != sys->nSelf) iluPattern(I, sys);
  if (I.factoredVersion == sys->version) return;
  parallelFor(I.nnz, IluFillKernel{I.src.p, sys->diag.p, sys->off.p, I.coef.p});
  for (size_t l = 0; l + 1 < I.startL.size(); l++)
    parallelFor(I.startL[l + 1] - I.startL[l],
                IluFactorRows{I.startL[l], I.orderL.p, I.row.p, I.col.p, I.diagIdx.p, I.coef.p});
  I.factoredVersion = sys->version;
}

// CRMatrix::iluSolve: x = U^-1 L^-1 (-b), i.e. A x + b = 0 approximately (F/CRMatrix.h:376-388)
static void iluApply(Ilu0& I, const double* b, double* x) {
  for (size_t l = 0; l + 1 < I.startL.size(); l++)
    parallelFor(I.startL[l + 1] - I.startL[l],
                IluLowerRows{I.startL[l], I.orderL.p, I.row.p, I.col.p, I.diagIdx.p, I.coef.p, b, I.y.p});
  for (size_t l = 0; l + 1 < I.startU.size(); l++)
    parallelFor(I.startU[l + 1] - I.startU[l],
                IluUpperRows{I.startU[l], I.orderU.p, I.row.p, I.col.p, I.diagIdx.p, I.coef.p, I.y.p, x});
}

Ilu0& Amg::iluFor(System* sys) {
  if (!ilu) ilu.reset(new Ilu0);
  iluEnsure(*ilu, sys);
  return *ilu;
}

// ILU0Solver::smooth on (b := rhs, delta := out), both in the system's natural numbering
void Amg::iluSmooth(System* sys, const double* rhs, double* out) { iluApply(iluFor(sys), rhs, out); }

int Amg::iluLevels(System* sys) { return (int)iluFor(sys).startL.size() - 1; }

// ILU0Solver::solve, F/ILU0Solver.cpp:46-93
void Amg::iluSolve(System* sys, int nMaxIterations, double relTol, double absTol, double* rnorm0Out, double* rnormOut,
                   int* itersOut) {
  requireReady();
  if (!scalars.p) scalars.alloc(16);
  history.clear();
  const bool multiRank = commActive() && sys->mesh && !sys->noHalo;
  Ilu0& I = iluFor(sys);
  const int n = sys->nSelf;
  auto norm = [&]() {
    reduceRows<1>(n, CsrResidualRows{sys->row, sys->col, sys->diag.p, sys->off.p, sys->b.p, sys->delta.p}, scalars.p + 8);
    if (multiRank) commAllreduceSum(scalars.p + 8, 1);
    double v;
    copyD2H(&v, scalars.p + 8, sizeof(double));
    return v;
  };
  const double rNorm0 = norm();
  history.push_back(rNorm0);
  double rNorm = rNorm0;
  int iters = 0;
  if (!(rNorm0 < absTol)) {
    for (int i = 1; i < nMaxIterations; i++) {
      iluApply(I, sys->b.p, sys->delta.p);
      if (multiRank) sys->mesh->halo.exchange(sys->delta.p, 1);
      iters++;
      rNorm = norm();
      history.push_back(rNorm);
      if (rNorm < absTol || rNorm / rNorm0 < relTol) break;
    }
  }
  totalIterations += iters;
  if (rnorm0Out) *rnorm0Out = rNorm0;
  if (rnormOut) *rnormOut = rNorm;
  if (itersOut) *itersOut = iters;
}

}  // namespace fvmgpu
