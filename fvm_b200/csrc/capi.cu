// fvm_b200 / libfvmgpu -- the extern "C" boundary declared in include/fvmgpu.h.
// Every entry point catches C++ exceptions and turns them into (non-zero return, message):
// the C-ABI equivalent of the reference's CException -> RuntimeError translation
// (F/CException.h:16-21, F/baseExt.i:49-59).
#include "solver.cuh"

using namespace fvmgpu;

static thread_local std::string g_lastError;

#define API_BEGIN try {
#define API_END                                  \
  return 0;                                      \
  }                                              \
  catch (const std::exception& e) {              \
    g_lastError = e.what();                      \
    return 1;                                    \
  }                                              \
  catch (...) {                                  \
    g_lastError = "unknown error";               \
    return 1;                                    \
  }

// the opaque handle types of fvmgpu.h are never defined: handles are the C++ objects' addresses
static Mesh* M(fvmgpu_mesh_t h) { if (!h) fail("null mesh handle"); return reinterpret_cast<Mesh*>(h); }
static System* S(fvmgpu_system_t h) { if (!h) fail("null system handle"); return reinterpret_cast<System*>(h); }
static Amg* A(fvmgpu_solver_t h) { if (!h) fail("null solver handle"); return reinterpret_cast<Amg*>(h); }
static Flow* FL(fvmgpu_flow_t h) { if (!h) fail("null flow handle"); return reinterpret_cast<Flow*>(h); }

extern "C" {

const char* fvmgpu_last_error(void) { return g_lastError.c_str(); }
int fvmgpu_version(void) { return FVMGPU_VERSION; }

void fvmgpu_amg_default_opts(fvmgpu_amg_opts* o) {
  // F/AMG.cpp:14-22 and F/LinearSolver.h:15-20
  o->nMaxIterations = 100;
  o->verbosity = 2;
  o->relativeTolerance = 1e-8;
  o->absoluteTolerance = 1e-50;
  o->maxCoarseLevels = 30;
  o->nPreSweeps = 0;
  o->nPostSweeps = 1;
  o->coarseGroupSize = 2;
  o->weightRatioThreshold = 0.65;
  o->cycleType = FVMGPU_CYCLE_V;
  o->smootherType = FVMGPU_SMOOTHER_GAUSS_SEIDEL;
}

int fvmgpu_init(int device) {
  API_BEGIN
  Context& c = ctx();
  if (c.ready) {
    if (c.device != device) fail("libfvmgpu already initialised on device %d", c.device);
    return 0;
  }
#ifndef FVMGPU_HOSTSIM
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    fail("libfvmgpu needs a CUDA device (sm_100a build, no CPU fallback): %s",
         e != cudaSuccess ? cudaGetErrorString(e) : "no device found");
  if (device < 0 || device >= count) fail("device %d out of range (have %d)", device, count);
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  c.smCount = prop.multiProcessorCount;
  CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  {
    cudaMemPool_t pool;
    CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ULL;  // never trim: freed blocks are reused by the next hierarchy
    CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  for (int i = 0; i < 16; i++) {
    CUDA_CHECK(cudaEventCreate(&c.timerStart[i]));
    CUDA_CHECK(cudaEventCreate(&c.timerStop[i]));
  }
  c.reduceScratch = (double*)devAlloc(sizeof(double) * 4 * kMaxReduceBlocks);
#else
  c.reduceScratch = (double*)devAlloc(sizeof(double) * 4 * kMaxReduceBlocks);
#endif
  c.device = device;
  c.ready = true;
  API_END
}

int fvmgpu_shutdown(void) {
  API_BEGIN
  Context& c = ctx();
  if (!c.ready) return 0;
  if (c.reduceScratch) devFree(c.reduceScratch);
  if (c.l2scratch) devFree(c.l2scratch);
  c.reduceScratch = nullptr;
  c.l2scratch = nullptr;
#ifndef FVMGPU_HOSTSIM
  devTrimCache();
  cudaStreamSynchronize(c.stream);
  for (int i = 0; i < 16; i++) { cudaEventDestroy(c.timerStart[i]); cudaEventDestroy(c.timerStop[i]); }
  cudaStreamDestroy(c.stream);
  c.stream = nullptr;
#endif
  c.ready = false;
  API_END
}

int fvmgpu_device_info(char* name, int cap, int* sm_count, double* mem_gb) {
  API_BEGIN
  requireReady();
#ifndef FVMGPU_HOSTSIM
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, ctx().device));
  if (name && cap > 0) { std::strncpy(name, prop.name, cap - 1); name[cap - 1] = 0; }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (mem_gb) *mem_gb = (double)prop.totalGlobalMem / 1e9;
#else
  if (name && cap > 0) { std::strncpy(name, "hostsim", cap - 1); name[cap - 1] = 0; }
  if (sm_count) *sm_count = 0;
  if (mem_gb) *mem_gb = 0;
#endif
  API_END
}

int fvmgpu_synchronize(void) {
  API_BEGIN
  requireReady();
  streamSync();
  API_END
}

int fvmgpu_timer_start(int slot) {
  API_BEGIN
  requireReady();
  if (slot < 0 || slot >= 16) fail("timer slot out of range");
#ifndef FVMGPU_HOSTSIM
  CUDA_CHECK(cudaEventRecord(ctx().timerStart[slot], ctx().stream));
#endif
  API_END
}
int fvmgpu_timer_stop(int slot, double* ms) {
  API_BEGIN
  requireReady();
  if (slot < 0 || slot >= 16) fail("timer slot out of range");
#ifndef FVMGPU_HOSTSIM
  CUDA_CHECK(cudaEventRecord(ctx().timerStop[slot], ctx().stream));
  CUDA_CHECK(cudaEventSynchronize(ctx().timerStop[slot]));
  float t = 0;
  CUDA_CHECK(cudaEventElapsedTime(&t, ctx().timerStart[slot], ctx().timerStop[slot]));
  if (ms) *ms = t;
#else
  if (ms) *ms = 0;
#endif
  API_END
}
int fvmgpu_counters(long long* kernel_launches, long long* h2d_bytes, long long* d2h_bytes) {
  if (kernel_launches) *kernel_launches = ctx().launches;
  if (h2d_bytes) *h2d_bytes = ctx().h2d;
  if (d2h_bytes) *d2h_bytes = ctx().d2h;
  return 0;
}
int fvmgpu_host_alloc(void** out, unsigned long long bytes) {
  API_BEGIN
  if (!out) fail("host_alloc: null output");
#ifdef FVMGPU_HOSTSIM
  *out = std::malloc(bytes ? bytes : 1);
#else
  requireReady();
  CUDA_CHECK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
#endif
  API_END
}
int fvmgpu_host_free(void* p) {
  API_BEGIN
#ifdef FVMGPU_HOSTSIM
  std::free(p);
#else
  if (p) CUDA_CHECK(cudaFreeHost(p));
#endif
  API_END
}
int fvmgpu_flush_l2(void) {
  API_BEGIN
  requireReady();
  const size_t bytes = 256u << 20;
  if (!ctx().l2scratch) ctx().l2scratch = devAlloc(bytes);
  devMemset(ctx().l2scratch, 1, bytes);
  API_END
}

int fvmgpu_profile_begin(void) {
  API_BEGIN
  requireReady();
  profileBegin();
  API_END
}
int fvmgpu_profile_end(int cap, char* names, int nameStride, long long* rows, long long* launches, double* ms,
                       int* count) {
  API_BEGIN
  requireReady();
  std::vector<ProfileRecord> recs = profileEnd();
  if (count) *count = (int)recs.size();
  for (int i = 0; i < (int)recs.size() && i < cap; i++) {
    std::string nm = recs[i].name;
    if (recs[i].tag >= 0) nm += "@L" + std::to_string(recs[i].tag);
    std::strncpy(names + (size_t)i * nameStride, nm.c_str(), nameStride - 1);
    names[(size_t)i * nameStride + nameStride - 1] = 0;
    rows[i] = recs[i].n;
    launches[i] = recs[i].launches;
    ms[i] = recs[i].ms;
  }
  API_END
}

// ---------------------------------------------------------------- mesh
int fvmgpu_mesh_create(fvmgpu_mesh_t* out, int dim, int nCellsSelf, int nCellsTotal, int nFaces,
                       const int* faceCells, const int* cellCellsRow, const int* cellCellsCol, int nGroups,
                       const int* groupOffset, const int* groupCount, const int* groupId, const int* groupKind) {
  API_BEGIN
  if (!out) fail("null output handle");
  *out = reinterpret_cast<fvmgpu_mesh_t>(meshCreate(dim, nCellsSelf, nCellsTotal, nFaces, faceCells, cellCellsRow,
                                                    cellCellsCol, nGroups, groupOffset, groupCount, groupId, groupKind));
  API_END
}
int fvmgpu_mesh_set_geometry(fvmgpu_mesh_t mesh, const double* faceArea, const double* faceAreaMag,
                             const double* faceCentroid, const double* cellCentroid, const double* cellVolume,
                             const int* ibType) {
  API_BEGIN
  if (!faceArea || !faceAreaMag || !cellCentroid || !cellVolume) fail("set_geometry: null array");
  meshSetGeometry(M(mesh), faceArea, faceAreaMag, faceCentroid, cellCentroid, cellVolume, ibType);
  API_END
}
int fvmgpu_mesh_compute_geometry(fvmgpu_mesh_t mesh, int nNodes, const double* nodeCoords, const int* faceNodeOffsets,
                                 const int* faceNodes, double* faceArea, double* faceAreaMag, double* faceCentroid,
                                 double* cellCentroid, double* cellVolume) {
  API_BEGIN
  if (!nodeCoords || !faceNodeOffsets || !faceNodes) fail("compute_geometry: null array");
  meshComputeGeometry(M(mesh), nNodes, nodeCoords, faceNodeOffsets, faceNodes, faceArea, faceAreaMag, faceCentroid,
                      cellCentroid, cellVolume);
  API_END
}
int fvmgpu_mesh_set_halo(fvmgpu_mesh_t mesh, int nNeigh, const int* peerRank, const int* scatterOff,
                         const int* scatterIdx, const int* gatherOff, const int* gatherIdx) {
  API_BEGIN
  meshSetHalo(M(mesh), nNeigh, peerRank, scatterOff, scatterIdx, gatherOff, gatherIdx);
  API_END
}
int fvmgpu_mesh_destroy(fvmgpu_mesh_t mesh) {
  API_BEGIN
  if (mesh) { streamSync(); delete M(mesh); }
  API_END
}
int fvmgpu_mesh_download_pair_to_col(fvmgpu_mesh_t mesh, int* pairToCol) {
  API_BEGIN
  Mesh* m = M(mesh);
  m->pairToCol.download(pairToCol, 2 * (size_t)m->nFaces);
  API_END
}
int fvmgpu_mesh_download_gradient_weights(fvmgpu_mesh_t mesh, double* coeffs) {
  API_BEGIN
  Mesh* m = M(mesh);
  if (!m->hasGeometry) fail("gradient weights: geometry not set");
  std::vector<double> w = m->gradW.toHost();
  const size_t nnz = (size_t)m->nnz;
  for (size_t k = 0; k < nnz; k++) {
    coeffs[3 * k] = w[k];
    coeffs[3 * k + 1] = w[nnz + k];
    coeffs[3 * k + 2] = w[2 * nnz + k];
  }
  API_END
}

// ---------------------------------------------------------------- system
int fvmgpu_system_create(fvmgpu_system_t* out, fvmgpu_mesh_t mesh) {
  API_BEGIN
  if (!out) fail("null output handle");
  *out = reinterpret_cast<fvmgpu_system_t>(systemCreate(M(mesh)));
  API_END
}
int fvmgpu_system_create_raw(fvmgpu_system_t* out, int nSelf, int nGhost, const int* row, const int* col,
                             const double* diag, const double* offdiag, const double* b) {
  API_BEGIN
  if (!out) fail("null output handle");
  if (nSelf <= 0 || nGhost < 0 || !row || !diag || !b) fail("system_create_raw: bad arguments");
  *out = reinterpret_cast<fvmgpu_system_t>(systemCreateRaw(nSelf, nGhost, row, col, diag, offdiag, b));
  API_END
}
int fvmgpu_system_destroy(fvmgpu_system_t sys) {
  API_BEGIN
  if (sys) { streamSync(); delete S(sys); }
  API_END
}
int fvmgpu_system_set_field(fvmgpu_system_t sys, int field, const double* host, long long n) {
  API_BEGIN
  if (!host) fail("set_field: null array");
  systemSetField(S(sys), field, host, n, false, 0.0);
  API_END
}
int fvmgpu_system_fill_field(fvmgpu_system_t sys, int field, double value) {
  API_BEGIN
  systemSetField(S(sys), field, nullptr, 0, true, value);
  API_END
}
int fvmgpu_system_get_field(fvmgpu_system_t sys, int field, double* host, long long n) {
  API_BEGIN
  if (!host) fail("get_field: null array");
  systemGetField(S(sys), field, host, n);
  API_END
}
int fvmgpu_system_set_bc(fvmgpu_system_t sys, int groupId, int bcKind, const double* p, int np,
                         const double* perFace) {
  API_BEGIN
  if (bcKind < 0 || bcKind > FVMGPU_BC_DIELECTRIC_INTERFACE) fail("set_bc: unknown BC kind %d", bcKind);
  systemSetBc(S(sys), groupId, bcKind, p, np, perFace);
  API_END
}
int fvmgpu_compute_gradient(fvmgpu_system_t sys) {
  API_BEGIN
  computeGradient(S(sys));
  API_END
}
int fvmgpu_assemble(fvmgpu_system_t sys, const fvmgpu_assemble_opts* opts) {
  API_BEGIN
  if (!opts) fail("assemble: null options");
  assemble(S(sys), *opts);
  API_END
}
int fvmgpu_download_system(fvmgpu_system_t sys, double* diag, double* offdiag, double* b, int* isBoundary) {
  API_BEGIN
  System* s = S(sys);
  if (diag) s->diag.download(diag, s->nTotal);
  if (offdiag) s->off.download(offdiag, (size_t)s->nnz);
  if (b) s->b.download(b, s->nTotal);
  if (isBoundary) s->isBoundary.download(isBoundary, s->nTotal);
  API_END
}

// ---------------------------------------------------------------- solvers
int fvmgpu_amg_create(fvmgpu_solver_t* out, const fvmgpu_amg_opts* opts) {
  API_BEGIN
  if (!out) fail("null output handle");
  Amg* a = new Amg;
  if (opts) a->opts = *opts;
  else fvmgpu_amg_default_opts(&a->opts);
  *out = reinterpret_cast<fvmgpu_solver_t>(a);
  API_END
}
int fvmgpu_amg_set_opts(fvmgpu_solver_t s, const fvmgpu_amg_opts* opts) {
  API_BEGIN
  if (!opts) fail("null options");
  Amg* a = A(s);
  const bool structural = a->opts.coarseGroupSize != opts->coarseGroupSize ||
                          a->opts.weightRatioThreshold != opts->weightRatioThreshold ||
                          a->opts.maxCoarseLevels != opts->maxCoarseLevels;
  a->opts = *opts;
  if (structural) a->cleanup();
  API_END
}
int fvmgpu_amg_solve(fvmgpu_solver_t s, fvmgpu_system_t sys, double* rnorm0, double* rnorm, int* iters) {
  API_BEGIN
  A(s)->solve(S(sys), rnorm0, rnorm, iters);
  API_END
}
int fvmgpu_amg_smooth(fvmgpu_solver_t s, fvmgpu_system_t sys) {
  API_BEGIN
  A(s)->smooth(S(sys));
  API_END
}
int fvmgpu_amg_cleanup(fvmgpu_solver_t s) {
  API_BEGIN
  streamSync();
  A(s)->cleanup();
  API_END
}
int fvmgpu_amg_destroy(fvmgpu_solver_t s) {
  API_BEGIN
  if (s) { streamSync(); delete A(s); }
  API_END
}
int fvmgpu_amg_levels(fvmgpu_solver_t s, int cap, int* nLevels, long long* sizes, long long* nnzs, int* colours) {
  API_BEGIN
  Amg* a = A(s);
  // distributed levels (this rank's rows), then the levels of the replicated merged hierarchy
  // below the merged level (multi-GPU; its level 0 is the merged level itself in global size)
  std::vector<Level*> all;
  for (auto& l : a->levels) all.push_back(l.get());
  if (a->nested) for (auto& l : a->nested->levels) all.push_back(l.get());
  const int nl = (int)all.size();
  if (nLevels) *nLevels = nl;
  for (int l = 0; l < nl && l < cap; l++) {
    if (sizes) sizes[l] = all[l]->n;
    if (nnzs) {
      if (all[l]->nnzTrue < 0 && all[l]->nnzDev.p) all[l]->nnzTrue = (long long)all[l]->nnzDev.hostAt(0);
      nnzs[l] = all[l]->nnzTrue;
    }
    if (colours) colours[l] = all[l]->nColours;
  }
  API_END
}
int fvmgpu_amg_level_col_bytes(fvmgpu_solver_t s, int cap, double* colBytes) {
  API_BEGIN
  Amg* a = A(s);
  std::vector<Level*> all;
  for (auto& l : a->levels) all.push_back(l.get());
  if (a->nested) for (auto& l : a->nested->levels) all.push_back(l.get());
  for (int l = 0; l < (int)all.size() && l < cap; l++) {
    Level& L = *all[l];
    double frac = 0.0;
    if (L.scol16.p && L.nSlices > 0) {
      if (L.compressedSlices < 0 && L.modeCount.p) L.compressedSlices = (long long)L.modeCount.hostAt(0);
      frac = (double)L.compressedSlices / (double)L.nSlices;
    }
    colBytes[l] = 4.0 - frac * (4.0 - 2.125);
  }
  API_END
}
int fvmgpu_amg_level_order(fvmgpu_solver_t s, int level, long long cap, int* nat, int* nColours,
                           long long* colourStart) {
  API_BEGIN
  Amg* a = A(s);
  if (level < 0 || level >= (int)a->levels.size()) fail("fvmgpu_amg_level_order: no such level");
  Level& L = *a->levels[level];
  if (nColours) *nColours = L.nColours;
  if (colourStart) for (int c = 0; c <= L.nColours; c++) colourStart[c] = L.colourStart[c];
  if (nat) {
    if (cap < L.n) fail("fvmgpu_amg_level_order: buffer too small");
    if (L.n) L.nat.download(nat, (size_t)L.n);
  }
  API_END
}
int fvmgpu_debug_tail_trace(int cap, unsigned long long* times_ns, int* tags, int* n) {
  API_BEGIN
  const int k = tailTraceRead(cap, times_ns, tags);
  if (n) *n = k;
  API_END
}
int fvmgpu_debug_set_aggregator(fvmgpu_aggregate_fn fn, void* user) {
  API_BEGIN
  setDebugAggregator(fn, user);
  API_END
}
int fvmgpu_amg_last_timing(fvmgpu_solver_t s, double* setup_ms, double* cycles_ms) {
  API_BEGIN
  Amg* a = A(s);
  if (setup_ms) *setup_ms = a->lastSetupMs;
  if (cycles_ms) *cycles_ms = a->lastCyclesMs;
  API_END
}
int fvmgpu_solver_history(fvmgpu_solver_t s, int cap, double* out, int* n) {
  API_BEGIN
  Amg* a = A(s);
  const int cnt = (int)a->history.size();
  if (n) *n = cnt;
  for (int i = 0; i < cnt && i < cap; i++) out[i] = a->history[i];
  API_END
}
int fvmgpu_bcgstab_solve(fvmgpu_solver_t precond, fvmgpu_system_t sys, int nMaxIterations, double relativeTolerance,
                         double absoluteTolerance, double* rnorm0, double* rnorm, int* iters) {
  API_BEGIN
  A(precond)->bcgstab(S(sys), nMaxIterations, relativeTolerance, absoluteTolerance, rnorm0, rnorm, iters);
  API_END
}
int fvmgpu_bcgstab_ilu0_solve(fvmgpu_solver_t s, fvmgpu_system_t sys, int nMaxIterations, double relativeTolerance,
                              double absoluteTolerance, double* rnorm0, double* rnorm, int* iters) {
  API_BEGIN
  Amg* a = A(s);
  // the Krylov vectors live on level 0 of a hierarchy; its coarse levels are not needed here
  const int keepLevels = a->opts.maxCoarseLevels, keepKind = a->precondKind;
  a->opts.maxCoarseLevels = 0;
  a->precondKind = 1;
  try {
    a->bcgstab(S(sys), nMaxIterations, relativeTolerance, absoluteTolerance, rnorm0, rnorm, iters);
  } catch (...) {
    a->opts.maxCoarseLevels = keepLevels; a->precondKind = keepKind;
    a->cleanup();
    throw;
  }
  a->opts.maxCoarseLevels = keepLevels; a->precondKind = keepKind;
  a->cleanup();
  API_END
}
int fvmgpu_ilu0_solve(fvmgpu_solver_t s, fvmgpu_system_t sys, int nMaxIterations, double relativeTolerance,
                      double absoluteTolerance, double* rnorm0, double* rnorm, int* iters, int* levels) {
  API_BEGIN
  A(s)->iluSolve(S(sys), nMaxIterations, relativeTolerance, absoluteTolerance, rnorm0, rnorm, iters);
  if (levels) *levels = A(s)->iluLevels(S(sys));
  API_END
}
int fvmgpu_cg_solve(fvmgpu_solver_t precond, fvmgpu_system_t sys, int nMaxIterations, double relativeTolerance,
                    double absoluteTolerance, double* rnorm0, double* rnorm, int* iters) {
  API_BEGIN
  A(precond)->cg(S(sys), nMaxIterations, relativeTolerance, absoluteTolerance, rnorm0, rnorm, iters);
  API_END
}
int fvmgpu_jacobi_solve(fvmgpu_solver_t s, fvmgpu_system_t sys, int nMaxIterations, double relativeTolerance,
                        double absoluteTolerance, double* rnorm0, double* rnorm, int* iters) {
  API_BEGIN
  A(s)->jacobiSolve(S(sys), nMaxIterations, relativeTolerance, absoluteTolerance, rnorm0, rnorm, iters);
  API_END
}
int fvmgpu_system_halo_exchange(fvmgpu_system_t sys, int field) {
  API_BEGIN
  systemHaloExchange(S(sys), field);
  API_END
}
int fvmgpu_post_solve_update(fvmgpu_system_t sys) {
  API_BEGIN
  postSolveUpdate(S(sys));
  API_END
}

// ---------------------------------------------------------------- ElectricModel (electric.cu)
int fvmgpu_electric_field(fvmgpu_system_t potential, double* E_host) {
  API_BEGIN
  electricField(S(potential), E_host);
  API_END
}
int fvmgpu_electric_drift_flux(fvmgpu_system_t potential, fvmgpu_system_t charge, double mobility, double vsat,
                               int nSymmetryGroups, const int* symmetryGroupIds, double* velocity_host) {
  API_BEGIN
  electricDriftFlux(S(potential), S(charge), mobility, vsat, nSymmetryGroups, symmetryGroupIds, velocity_host);
  API_END
}

// ---------------------------------------------------------------- FlowModel (flow.cu)
int fvmgpu_flow_create(fvmgpu_flow_t* out, fvmgpu_mesh_t mesh) {
  API_BEGIN
  if (!out) fail("null output handle");
  *out = reinterpret_cast<fvmgpu_flow_t>(flowCreate(M(mesh)));
  API_END
}
int fvmgpu_flow_destroy(fvmgpu_flow_t flow) {
  API_BEGIN
  if (flow) flowDestroy(FL(flow));
  API_END
}
int fvmgpu_flow_set_field(fvmgpu_flow_t flow, int field, const double* host, long long n) {
  API_BEGIN
  if (!host) fail("flow_set_field: null array");
  flowSetField(FL(flow), field, host, n, false, 0.0);
  API_END
}
int fvmgpu_flow_fill_field(fvmgpu_flow_t flow, int field, double value) {
  API_BEGIN
  flowSetField(FL(flow), field, nullptr, 0, true, value);
  API_END
}
int fvmgpu_flow_get_field(fvmgpu_flow_t flow, int field, double* host, long long n) {
  API_BEGIN
  if (!host) fail("flow_get_field: null array");
  flowGetField(FL(flow), field, host, n);
  API_END
}
int fvmgpu_flow_set_reference_cell(fvmgpu_flow_t flow, int localCell) {
  API_BEGIN
  flowSetReferenceCell(FL(flow), localCell);
  API_END
}
int fvmgpu_flow_set_bc(fvmgpu_flow_t flow, int groupId, int bcKind, const double* p, int np) {
  API_BEGIN
  flowSetBc(FL(flow), groupId, bcKind, p, np);
  API_END
}
int fvmgpu_flow_init(fvmgpu_flow_t flow) {
  API_BEGIN
  flowInit(FL(flow));
  API_END
}
int fvmgpu_flow_assemble_momentum(fvmgpu_flow_t flow, const fvmgpu_flow_opts* opts) {
  API_BEGIN
  if (!opts) fail("null options");
  flowAssembleMomentum(FL(flow), *opts);
  API_END
}
int fvmgpu_flow_download_momentum(fvmgpu_flow_t flow, double* diag3, double* offdiag, double* b3) {
  API_BEGIN
  flowDownloadMomentum(FL(flow), diag3, offdiag, b3);
  API_END
}
int fvmgpu_flow_solve_momentum(fvmgpu_flow_t flow, fvmgpu_solver_t solver, int bcgstab, int bcgMaxIterations,
                               double bcgRelTol, double bcgAbsTol, double* rnorm0, int* iters) {
  API_BEGIN
  flowSolveMomentum(FL(flow), A(solver), bcgstab, bcgMaxIterations, bcgRelTol, bcgAbsTol, rnorm0, iters);
  API_END
}
int fvmgpu_flow_assemble_continuity(fvmgpu_flow_t flow, const fvmgpu_flow_opts* opts) {
  API_BEGIN
  if (!opts) fail("null options");
  flowAssembleContinuity(FL(flow), *opts);
  API_END
}
int fvmgpu_flow_download_continuity(fvmgpu_flow_t flow, double* diag, double* offdiag, double* b, int* isBoundary) {
  API_BEGIN
  flowDownloadContinuity(FL(flow), diag, offdiag, b, isBoundary);
  API_END
}
int fvmgpu_flow_solve_continuity(fvmgpu_flow_t flow, fvmgpu_solver_t solver, int bcgstab, int bcgMaxIterations,
                                 double bcgRelTol, double bcgAbsTol, const fvmgpu_flow_opts* opts, double* rnorm0,
                                 int* iters) {
  API_BEGIN
  if (!opts) fail("null options");
  flowSolveContinuity(FL(flow), A(solver), bcgstab, bcgMaxIterations, bcgRelTol, bcgAbsTol, *opts, rnorm0, iters);
  API_END
}

// ---------------------------------------------------------------- multi-GPU plumbing (comm.cu)
int fvmgpu_comm_unique_id(void* out128) {
  API_BEGIN
  commUniqueId(out128);
  API_END
}
int fvmgpu_comm_init(int nranks, int rank, const void* uniqueId128) {
  API_BEGIN
  requireReady();
  if (nranks < 1 || rank < 0 || rank >= nranks) fail("comm_init: bad rank %d of %d", rank, nranks);
  commInitNccl(nranks, rank, uniqueId128);
  ctx().nranks = nranks;
  ctx().rank = rank;
  peerInit();   // NVLink peer-memory transport on top (collective); stays inactive where it cannot be set up
  API_END
}
int fvmgpu_comm_destroy(void) {
  API_BEGIN
  commDestroy();
  ctx().nranks = 1;
  ctx().rank = 0;
  API_END
}
int fvmgpu_comm_counters(long long* collectives) {
  API_BEGIN
  if (collectives) *collectives = ctx().collectives;
  API_END
}
#ifdef FVMGPU_HOSTSIM
int fvmgpu_hostsim_set_comm(fvmgpu_hostsim_exchange_fn e, fvmgpu_hostsim_allreduce_fn r, fvmgpu_hostsim_allgather_fn g) {
  API_BEGIN
  hostsimSetComm(e, r, g);
  API_END
}
#endif

}  // extern "C"
