// fvm_b200 / libfvmgpu -- context, raw memory, scan / sort primitives.
// Device build: CUDA runtime + CUB (device-wide scan / radix sort used by the SETUP phases only).
// FVMGPU_HOSTSIM build (tests only, see common.cuh): malloc / std algorithms.
#include "common.cuh"

#ifndef FVMGPU_HOSTSIM
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>
#include <unordered_map>
#else
#include <algorithm>
#include <numeric>
#endif

namespace fvmgpu {

Context& ctx() {
  static Context c;
  return c;
}

unsigned long long nextVersion() {
  static unsigned long long counter = 0;  // one caller thread per process (include/fvmgpu.h)
  return ++counter;
}

void requireReady() {
  if (!ctx().ready) fail("libfvmgpu: not initialised (call fvmgpu_init; a CUDA device is required, there is no CPU path)");
}

#ifndef FVMGPU_HOSTSIM
// Stream-ordered allocation from the device's default memory pool (release threshold = never):
// the hierarchy is rebuilt every outer iteration (F/ThermalModel_impl.h:428,446), and a
// cudaMalloc/cudaFree pair per buffer per iteration would cost more than the setup kernels.
// FVMGPU_POISON=1 (debug): fill every new allocation with 0xFF bytes (NaN doubles, -1 ints) so that a
// read of memory the library never wrote shows up as NaN / a fault instead of depending on what the
// pool handed back
static bool poisonAllocs() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FVMGPU_POISON"); on = (e && atoi(e)) ? 1 : 0; }
  return on == 1;
}
// FVMGPU_TRACE_ALLOC=1 (debug): report allocator calls that block the host for more than 1 ms
static bool traceAllocs() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FVMGPU_TRACE_ALLOC"); on = (e && atoi(e)) ? 1 : 0; }
  return on == 1;
}
// Block cache in front of the pool. The hierarchy is rebuilt every outer iteration with the same
// sequence of buffer sizes, but the pool cannot always serve an 800 MB request from its free list and
// then maps fresh physical memory -- measured: single cudaMallocAsync calls blocking the host for
// 100-770 ms inside a 640 ms step. Freed blocks of >= 1 MB are therefore kept here, keyed by size, and
// handed back to the next request of (nearly) that size; everything runs on the one context stream, so
// reuse is stream-ordered like the pool's own. The cache is capped (FVMGPU_CACHE_GB, default a quarter
// of the device memory): past the cap it is released to the pool wholesale.
namespace {
struct BlockCache {
  std::mutex mu;
  std::multimap<size_t, void*> freeBlocks;
  std::unordered_map<void*, size_t> live;  // big blocks handed out
  size_t cachedBytes = 0, capBytes = 0;
};
BlockCache& blockCache() { static BlockCache c; return c; }
constexpr size_t kCacheMinBytes = 1u << 20;
}  // namespace
void devTrimCache() {
  BlockCache& c = blockCache();
  std::lock_guard<std::mutex> lock(c.mu);
  for (auto& kv : c.freeBlocks) {
    if (ctx().stream) cudaFreeAsync(kv.second, ctx().stream);
    else cudaFree(kv.second);
  }
  c.freeBlocks.clear();
  c.cachedBytes = 0;
}
void* devAlloc(size_t bytes) {
  void* p = nullptr;
  const bool big = bytes >= kCacheMinBytes && ctx().stream;
  if (big) {
    BlockCache& c = blockCache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.freeBlocks.lower_bound(bytes);
    if (it != c.freeBlocks.end() && it->first <= bytes + bytes / 8) {
      p = it->second;
      c.live[p] = it->first;
      c.cachedBytes -= it->first;
      c.freeBlocks.erase(it);
    }
  }
  if (!p) {
    const auto t0 = std::chrono::steady_clock::now();
    if (ctx().stream) {
      cudaError_t e = cudaMallocAsync(&p, bytes, ctx().stream);
      if (e == cudaErrorMemoryAllocation) {
        // the cached blocks count as used memory: hand them back and try once more
        (void)cudaGetLastError();
        devTrimCache();
        cudaStreamSynchronize(ctx().stream);
        e = cudaMallocAsync(&p, bytes, ctx().stream);
      }
      CUDA_CHECK(e);
    } else {
      CUDA_CHECK(cudaMalloc(&p, bytes));
    }
    if (traceAllocs()) {
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms > 1.0) fprintf(stderr, "[fvmgpu] alloc of %zu bytes blocked the host for %.2f ms\n", bytes, ms);
    }
    if (big) {
      BlockCache& c = blockCache();
      std::lock_guard<std::mutex> lock(c.mu);
      c.live[p] = bytes;
    }
  }
  if (poisonAllocs() && bytes) CUDA_CHECK(cudaMemsetAsync(p, 0xff, bytes, ctx().stream));
  return p;
}
void devFree(void* p) {
  if (!p) return;
  bool trim = false;
  {
    BlockCache& c = blockCache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.live.find(p);
    if (it != c.live.end() && ctx().stream) {
      if (!c.capBytes) {
        size_t freeB = 0, totalB = 0;
        double gb = 0;
        if (const char* e = getenv("FVMGPU_CACHE_GB")) gb = atof(e);
        if (gb > 0) c.capBytes = (size_t)(gb * 1e9);
        else if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) c.capBytes = totalB / 4;
        else c.capBytes = (size_t)8 << 30;
      }
      c.freeBlocks.emplace(it->second, p);
      c.cachedBytes += it->second;
      c.live.erase(it);
      trim = c.cachedBytes > c.capBytes;
      p = nullptr;
    } else if (it != c.live.end()) {
      c.live.erase(it);
    }
  }
  if (trim) devTrimCache();
  if (!p) return;
  if (ctx().stream) cudaFreeAsync(p, ctx().stream);
  else cudaFree(p);
}
void devMemset(void* p, int byte, size_t bytes) { CUDA_CHECK(cudaMemsetAsync(p, byte, bytes, ctx().stream)); }
void copyH2D(void* d, const void* h, size_t bytes) {
  CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx().stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx().stream));  // pageable host source may be reused by the caller
  ctx().h2d += (long long)bytes;
}
void copyD2H(void* h, const void* d, size_t bytes) {
  CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx().stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
  ctx().d2h += (long long)bytes;
}
void copyD2D(void* d, const void* s, size_t bytes) {
  CUDA_CHECK(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, ctx().stream));
}
void streamSync() { CUDA_CHECK(cudaStreamSynchronize(ctx().stream)); }

static DBuf<char>& cubTemp() {
  static DBuf<char> t;
  return t;
}

void exclusiveScan(const int* in_d, int* out_d, long long n) {
  // out has n+1 entries: out[n] = total. Scan n+1 items where the extra input is ignored:
  // ExclusiveSum over n+1 outputs needs n+1 inputs; use a temp copy when in aliases out.
  if (n < 0) return;
  DBuf<int> tmp((size_t)n + 1);
  copyD2D(tmp.p, in_d, (size_t)n * sizeof(int));
  devMemset(tmp.p + n, 0, sizeof(int));
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, tmp.p, out_d, (int)(n + 1), ctx().stream));
  cubTemp().ensure(bytes);
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(cubTemp().p, bytes, tmp.p, out_d, (int)(n + 1), ctx().stream));
  ctx().launches += 2;
  // (no synchronisation: tmp goes back to the stream-ordered pool / block cache, whose reuse is ordered on this stream)
}

void sortPairs(int* keys_d, int* vals_d, long long n, int bits) {
  if (n <= 0) return;
  DBuf<int> k2((size_t)n), v2((size_t)n);
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_d, k2.p, vals_d, v2.p, (int)n, 0, bits, ctx().stream));
  cubTemp().ensure(bytes);
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cubTemp().p, bytes, keys_d, k2.p, vals_d, v2.p, (int)n, 0, bits,
                                             ctx().stream));
  copyD2D(keys_d, k2.p, (size_t)n * sizeof(int));
  copyD2D(vals_d, v2.p, (size_t)n * sizeof(int));
  ctx().launches += 4;
}

// ---- profiler
namespace {
struct ProfEntry { const char* name; long long n; cudaEvent_t a, b; int tag; };
std::vector<ProfEntry>& profEntries() { static std::vector<ProfEntry> v; return v; }
std::vector<cudaEvent_t>& profPool() { static std::vector<cudaEvent_t> v; return v; }
cudaEvent_t profEvent() {
  auto& pool = profPool();
  if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
  cudaEvent_t e;
  CUDA_CHECK(cudaEventCreate(&e));
  return e;
}
}  // namespace
ProfileScope::ProfileScope(const char* name, long long n) : on(ctx().profiling) {
  if (!on) return;
  ProfEntry e{name, n, profEvent(), profEvent(), ctx().profileTag};
  cudaEventRecord(e.a, ctx().stream);
  profEntries().push_back(e);
}
ProfileScope::~ProfileScope() {
  if (on) cudaEventRecord(profEntries().back().b, ctx().stream);
}
void profileBegin() {
  streamSync();
  profEntries().clear();
  ctx().profiling = true;
}
std::vector<ProfileRecord> profileEnd() {
  ctx().profiling = false;
  streamSync();
  std::vector<ProfileRecord> out;
  for (ProfEntry& e : profEntries()) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e.a, e.b);
    bool found = false;
    for (ProfileRecord& r : out)
      if (r.n == e.n && r.tag == e.tag && r.name == e.name) { r.launches++; r.ms += ms; found = true; break; }
    if (!found) out.push_back(ProfileRecord{e.name, e.n, 1, (double)ms, e.tag});
    profPool().push_back(e.a);
    profPool().push_back(e.b);
  }
  profEntries().clear();
  return out;
}
#else
ProfileScope::ProfileScope(const char*, long long) : on(false) {}
ProfileScope::~ProfileScope() {}
void profileBegin() {}
std::vector<ProfileRecord> profileEnd() { return {}; }
void* devAlloc(size_t bytes) {
  void* p = std::malloc(bytes ? bytes : 1);
  static int poison = -1;
  if (poison < 0) { const char* e = getenv("FVMGPU_POISON"); poison = (e && atoi(e)) ? 1 : 0; }
  if (poison == 1 && p) std::memset(p, 0xff, bytes);
  return p;
}
void devFree(void* p) { std::free(p); }
void devMemset(void* p, int byte, size_t bytes) { std::memset(p, byte, bytes); }
void copyH2D(void* d, const void* h, size_t bytes) { std::memcpy(d, h, bytes); ctx().h2d += (long long)bytes; }
void copyD2H(void* h, const void* d, size_t bytes) { std::memcpy(h, d, bytes); ctx().d2h += (long long)bytes; }
void copyD2D(void* d, const void* s, size_t bytes) { std::memmove(d, s, bytes); }
void streamSync() {}
void exclusiveScan(const int* in_d, int* out_d, long long n) {
  int acc = 0;
  for (long long i = 0; i < n; i++) { int v = in_d[i]; out_d[i] = acc; acc += v; }
  out_d[n] = acc;
}
void sortPairs(int* keys_d, int* vals_d, long long n, int) {
  std::vector<std::pair<int, int>> v((size_t)n);
  for (long long i = 0; i < n; i++) v[(size_t)i] = {keys_d[i], vals_d[i]};
  std::stable_sort(v.begin(), v.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
  for (long long i = 0; i < n; i++) { keys_d[i] = v[(size_t)i].first; vals_d[i] = v[(size_t)i].second; }
}
#endif

}  // namespace fvmgpu
