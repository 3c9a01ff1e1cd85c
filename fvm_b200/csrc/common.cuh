// fvm_b200 / libfvmgpu -- shared plumbing: error capture, device buffers, the row-parallel
// launcher every kernel of the library goes through, deterministic reductions and scans.
//
// B200 (sm_100a) only; the shipped library (nvcc build) has NO CPU path.
// FVMGPU_HOSTSIM is a TEST-ONLY build mode (tests/hostsim): the same kernel functors are run by a
// single-threaded loop so that index logic can be unit-tested in the GPU-less dev container. It
// is never compiled into libfvmgpu.so, never loaded by fvm_b200/ and never timed.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <typeinfo>
#include <string>
#include <vector>

#ifdef FVMGPU_HOSTSIM
#include <cmath>
#define FVM_DEV inline
#define FVM_RESTRICT
struct double4 { double x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
inline int atomicCAS(int* a, int cmp, int v) { int o = *a; if (o == cmp) *a = v; return o; }
inline int atomicOr(int* a, int v) { int o = *a; *a |= v; return o; }
inline int atomicAdd(int* a, int v) { int o = *a; *a += v; return o; }
inline int atomicMax(int* a, int v) { int o = *a; if (v > o) *a = v; return o; }
inline int atomicMin(int* a, int v) { int o = *a; if (v < o) *a = v; return o; }
inline double atomicAdd(double* a, double v) { double o = *a; *a += v; return o; }
typedef int cudaStream_t_sim;
#else
#include <cuda_runtime.h>
#define FVM_DEV __device__ __forceinline__
#define FVM_RESTRICT __restrict__
#endif

namespace fvmgpu {

// ---------------------------------------------------------------- errors
struct Error : std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};
[[noreturn]] inline void fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error(buf);
}

#ifndef FVMGPU_HOSTSIM
#define CUDA_CHECK(expr)                                                                      \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      ::fvmgpu::fail("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
                     cudaGetErrorString(e__));                                                \
  } while (0)
#endif

// ---------------------------------------------------------------- global context
struct Context {
  bool ready = false;
  bool profiling = false;
  int profileTag = -1;  // AMG level of the launches being issued (-1: not inside the solver); suffix '@L<k>' of profile names
  int device = -1;
  int smCount = 148;
#ifndef FVMGPU_HOSTSIM
  cudaStream_t stream = nullptr;  // compute stream: every kernel of the library runs here
  cudaEvent_t timerStart[16] = {};
  cudaEvent_t timerStop[16] = {};
#endif
  long long launches = 0;  // kernels launched by this library
  long long h2d = 0, d2h = 0;
  long long collectives = 0;  // halo exchanges / all-reduces / all-gathers issued
  void* l2scratch = nullptr;
  double* reduceScratch = nullptr;  // per-block partial sums (deterministic two-stage reductions)
  double* reduceHost = nullptr;     // pinned host landing zone for reduction results
  // multi-GPU
  int nranks = 1, rank = 0;
  void* ncclComm = nullptr;
};
Context& ctx();
void requireReady();
// Process-wide, strictly increasing stamp: every System gets a fresh one when it is created and whenever its
// matrix (diag / off-diagonal / pattern) changes. Cached hierarchies and ILU factors are keyed on the stamp
// alone -- an address can be handed out again by malloc / the block cache, a stamp cannot.
unsigned long long nextVersion();

inline int ceilDiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- raw device memory
void* devAlloc(size_t bytes);
void devFree(void* p);
void devTrimCache();  // device build: hand the cached free blocks back to the CUDA pool
void devMemset(void* p, int byte, size_t bytes);          // stream ordered
void copyH2D(void* d, const void* h, size_t bytes);       // stream ordered (pageable source: staged by the driver)
void copyD2H(void* h, const void* d, size_t bytes);       // stream ordered + synchronises
void copyD2D(void* d, const void* s, size_t bytes);       // stream ordered
void streamSync();

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() {}
  explicit DBuf(size_t n_) { alloc(n_); }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void release() {
    if (p) devFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t n_) {
    release();
    n = n_;
    if (n) p = (T*)devAlloc(n * sizeof(T));
  }
  void ensure(size_t n_) { if (n_ > n) alloc(n_); }
  void zero() { if (n) devMemset(p, 0, n * sizeof(T)); }
  void fillBytes(int byte) { if (n) devMemset(p, byte, n * sizeof(T)); }
  void upload(const T* h, size_t cnt) {
    if (cnt > n) alloc(cnt);
    if (cnt) copyH2D(p, h, cnt * sizeof(T));
  }
  void download(T* h, size_t cnt) const {
    if (cnt > n) fail("download of %zu elements from a buffer of %zu", cnt, n);
    if (cnt) copyD2H(h, p, cnt * sizeof(T));
  }
  std::vector<T> toHost() const {
    std::vector<T> v(n);
    download(v.data(), n);
    return v;
  }
  T hostAt(size_t i) const {
    T v;
    copyD2H(&v, p + i, sizeof(T));
    return v;
  }
};

// ---------------------------------------------------------------- per-launch profiler
// When enabled (fvmgpu_profile_begin) every launch is bracketed by two CUDA events on the compute
// stream; fvmgpu_profile_end resolves them into (kernel class, rows, launches, total ms) records.
// Used by bench.py for the live roofline figure; off by default (zero overhead: one branch).
struct ProfileScope {
  bool on;
  ProfileScope(const char* name, long long n);
  ~ProfileScope();
};
void profileBegin();
struct ProfileRecord { std::string name; long long n; long long launches; double ms; int tag; };
std::vector<ProfileRecord> profileEnd();

// ---------------------------------------------------------------- row-parallel launcher
// Every kernel of the library is a functor with `void operator()(long long i) const`, one
// logical thread per row / face / entry; consecutive i map to consecutive lanes, so the SoA /
// SELL layouts give coalesced access. Grid = ceil(n/256) CTAs of 256 threads.
#ifdef FVMGPU_HOSTSIM
template <class F>
void parallelFor(long long n, const F& f) {
  for (long long i = 0; i < n; i++) f(i);
  ctx().launches++;
}
#else
template <class F>
__global__ void __launch_bounds__(256) k_rows(long long n, const F f) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f(i);
}
template <class F>
void parallelFor(long long n, const F& f) {
  if (n <= 0) return;
  ProfileScope prof(typeid(F).name(), n);
  k_rows<F><<<ceilDiv(n, 256), 256, 0, ctx().stream>>>(n, f);
  ctx().launches++;
  CUDA_CHECK(cudaGetLastError());
}
#endif

// Deterministic sum reduction of up to 4 values per row: f(i, double v[NV]) fills v.
// Stage 1: per-CTA sums in a fixed tree order -> scratch; stage 2: one CTA adds the partials in
// index order. Result lands in out_d[0..NV-1] (device). No atomics: bit-reproducible run to run.
constexpr int kReduceBlock = 256;
// One row per thread up to 9.7 M rows (a level-0 colour class of the 256^3 mesh is 8.4 M): a row of the fused
// residual + norm is a chain of dependent gathers, and a thread that walks several rows one after the other keeps
// fewer of them in flight than the one-row-per-thread smoother does (0.66 vs 0.89 of the HBM roof with 4.6 rows per
// thread). Earlier history: a 1.3-wave grid left half the machine idle in the tail (4.3 TB/s).
constexpr int kMaxReduceBlocks = 148 * 256;
#ifdef FVMGPU_HOSTSIM
template <int NV, class F>
void reduceRows(long long n, const F& f, double* out_d) {
  double acc[NV];
  for (int k = 0; k < NV; k++) acc[k] = 0;
  for (long long i = 0; i < n; i++) {
    double v[NV];
    for (int k = 0; k < NV; k++) v[k] = 0;
    f(i, v);
    for (int k = 0; k < NV; k++) acc[k] += v[k];
  }
  for (int k = 0; k < NV; k++) out_d[k] = acc[k];
  ctx().launches += 2;
}
#else
template <int NV, class F>
__global__ void __launch_bounds__(kReduceBlock) k_reduce1(long long n, const F f, double* partial) {
  __shared__ double sm[NV][kReduceBlock / 32];
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) acc[k] = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double v[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = 0.0;
    f(i, v);
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] += v[k];
  }
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double v = 0.0;
    for (int w = 0; w < kReduceBlock / 32; w++) v += sm[threadIdx.x][w];
    partial[(size_t)blockIdx.x * NV + threadIdx.x] = v;
  }
}
template <int NV>
__global__ void __launch_bounds__(256) k_reduce2(int nBlocks, const double* partial, double* out) {
  // one CTA per value; threads stride over the partials, then fixed shuffle / shared-memory trees
  __shared__ double sm[8];
  const int k = blockIdx.x, lane = threadIdx.x & 31;
  double v = 0.0;
  for (int b = threadIdx.x; b < nBlocks; b += 256) v += partial[(size_t)b * NV + k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += sm[w];
    out[k] = s;
  }
}
template <int NV, class F>
void reduceRows(long long n, const F& f, double* out_d) {
  int nb = ceilDiv(n, kReduceBlock);
  if (nb > kMaxReduceBlocks) nb = kMaxReduceBlocks;
  if (nb < 1) nb = 1;
  ProfileScope prof(typeid(F).name(), n);
  k_reduce1<NV, F><<<nb, kReduceBlock, 0, ctx().stream>>>(n, f, ctx().reduceScratch);
  k_reduce2<NV><<<NV, 256, 0, ctx().stream>>>(nb, ctx().reduceScratch, out_d);
  ctx().launches += 2;
  CUDA_CHECK(cudaGetLastError());
}
#endif

// exclusive prefix sum of n ints (in -> out, out[n] = total); in may alias out
void exclusiveScan(const int* in_d, int* out_d, long long n);
// stable sort of (key,value) int pairs by key, keys in [0, 2^bits)
void sortPairs(int* keys_d, int* vals_d, long long n, int bits);

}  // namespace fvmgpu
