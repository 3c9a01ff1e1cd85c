// tools/ipc_probe.cu -- does CUDA IPC peer memory work between two processes on this box, and what do a
// device-side flag round trip and a peer-store "push" cost over NVLink?  (measurement tool, not product)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ipc_probe tools/ipc_probe.cu && gpurun_out/ipc_probe
// Two processes (fork), one GPU each. Each allocates a buffer, exports it with cudaIpcGetMemHandle, opens the
// other's, then: (1) ping-pong of a flag written with st.release.sys into the peer's memory, (2) push of N
// doubles into the peer's buffer + flag, timed with CUDA events around a kernel that does push + signal + wait.
#include <cuda_runtime.h>
#include <sys/wait.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(e) do { cudaError_t r__ = (e); if (r__ != cudaSuccess) { fprintf(stderr, "[%d] CUDA %s at line %d: %s\n", me, #e, __LINE__, cudaGetErrorString(r__)); exit(2); } } while (0)

__device__ __forceinline__ void stRelease(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ldAcquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ int g_timeouts = 0;
// spin until *p >= v, giving up after ~1 s (so that a broken transport shows up as a count, not as a hung box)
__device__ __forceinline__ void waitFlag(const unsigned long long* p, unsigned long long v) {
  if (*(volatile int*)&g_timeouts) return;   // sticky: one failure ends all waiting
  const long long t0 = clock64();
  while (ldAcquire(p) < v) {
    if (clock64() - t0 > 2000000000LL) { atomicAdd(&g_timeouts, 1); return; }
  }
}

// rank 0 sends k, waits for the echo k; rank 1 waits for k, echoes. `iters` round trips in one kernel.
__global__ void pingpong(int me, unsigned long long* mine, unsigned long long* theirs, int iters, unsigned long long base) {
  for (int k = 1; k <= iters; k++) {
    const unsigned long long v = base + k;
    if (me == 0) {
      stRelease(theirs, v);
      waitFlag(mine, v);
    } else {
      waitFlag(mine, v);
      stRelease(theirs, v);
    }
  }
}

// push n doubles into the peer's buffer, last CTA signals, everybody waits for the peer's signal
__global__ void pushExchange(const double* src, double* peerDst, long long n, unsigned* done, unsigned long long* myFlag,
                             unsigned long long* peerFlag, unsigned long long epoch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) peerDst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(done, 1u);
    if (old == gridDim.x - 1) { *done = 0; stRelease(peerFlag, epoch); }
    waitFlag(myFlag, epoch);
  }
  __syncthreads();
}

// push-only variants (no wait): bytes per store instruction and stores in flight per thread
__global__ void pushV1(const double* src, double* dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
  __threadfence_system();
}
__global__ void pushV2(const double2* src, double2* dst, long long n2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2) dst[i] = src[i];
  __threadfence_system();
}
__global__ void pushV2x4(const double2* src, double2* dst, long long n2) {   // 4 x 16 B per thread, grid-stride
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double2 v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) if (i + k * stride < n2) v[k] = src[i + k * stride];
#pragma unroll
  for (int k = 0; k < 4; k++) if (i + k * stride < n2) dst[i + k * stride] = v[k];
  __threadfence_system();
}
__global__ void pushNoFence(const double* src, double* dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

#define STAGE(msg) fprintf(stderr, "[%d] %s\n", me, msg)
int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int p2c[2], c2p[2];
  if (pipe(p2c) || pipe(c2p)) return 1;
  const pid_t pid = fork();
  const int me = pid == 0 ? 1 : 0;
  const int rd = me == 0 ? c2p[0] : p2c[0], wr = me == 0 ? p2c[1] : c2p[1];
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { if (me == 0) printf("ipc_probe: needs 2 GPUs (found %d)\n", ndev); return 0; }
  CK(cudaSetDevice(me));
  { int can = 0; CK(cudaDeviceCanAccessPeer(&can, me, 1 - me)); fprintf(stderr, "[%d] canAccessPeer=%d\n", me, can); }
  const size_t nMax = 1 << 22;  // doubles
  char* buf = nullptr;
  const size_t bytes = nMax * 8 + 4096;
  CK(cudaMalloc(&buf, bytes));
  CK(cudaMemset(buf, 0, bytes));
  STAGE("allocated");
  cudaIpcMemHandle_t h, ho;
  CK(cudaIpcGetMemHandle(&h, buf));
  if (write(wr, &h, sizeof(h)) != (ssize_t)sizeof(h)) return 1;
  if (read(rd, &ho, sizeof(ho)) != (ssize_t)sizeof(ho)) return 1;
  STAGE("handles exchanged");
  char* peer = nullptr;
  CK(cudaIpcOpenMemHandle((void**)&peer, ho, cudaIpcMemLazyEnablePeerAccess));
  if (me == 0) printf("ipc_probe: cudaIpcOpenMemHandle OK (peer buffer mapped)\n");
  unsigned long long* myFlag = (unsigned long long*)buf;
  unsigned long long* peerFlag = (unsigned long long*)peer;
  unsigned long long* myFlag2 = myFlag + 16;
  unsigned long long* peerFlag2 = peerFlag + 16;
  double* myData = (double*)(buf + 4096);
  double* peerData = (double*)(peer + 4096);
  double* src = nullptr;
  unsigned* done = nullptr;
  CK(cudaMalloc(&src, nMax * 8));
  CK(cudaMemset(src, 1, nMax * 8));
  CK(cudaMalloc(&done, 4));
  CK(cudaMemset(done, 0, 4));
  CK(cudaDeviceSynchronize());
  char c = 1;
  if (write(wr, &c, 1) != 1 || read(rd, &c, 1) != 1) return 1;  // both mapped
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  // (0) plain peer memcpy through the mapping
  { double v = 3.25 + me; CK(cudaMemcpy(peer + 2048, &v, 8, cudaMemcpyHostToDevice)); CK(cudaDeviceSynchronize());
    if (write(wr, &c, 1) != 1 || read(rd, &c, 1) != 1) return 1;
    double got = 0; CK(cudaMemcpy(&got, buf + 2048, 8, cudaMemcpyDeviceToHost));
    fprintf(stderr, "[%d] peer memcpy landed: %s\n", me, got == 3.25 + (1 - me) ? "yes" : "NO"); }
  STAGE("ping-pong warm-up");
  // (1) flag ping-pong
  pingpong<<<1, 1>>>(me, myFlag, peerFlag, 100, 0);
  CK(cudaDeviceSynchronize());
  { int to = 0; CK(cudaMemcpyFromSymbol(&to, g_timeouts, 4)); fprintf(stderr, "[%d] warm-up done, timeouts=%d\n", me, to); if (to) { fprintf(stderr, "[%d] device-side flags over IPC do NOT work here\n", me); return 3; } }
  const int iters = 2000;
  CK(cudaEventRecord(a));
  pingpong<<<1, 1>>>(me, myFlag, peerFlag, iters, 1000);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  if (me == 0) printf("ipc_probe: flag round trip %.2f us (one way ~%.2f us)\n", ms * 1e3 / iters, ms * 1e3 / iters / 2);
  // (2) push + signal + wait, one kernel per exchange, back to back
  unsigned long long epoch = 0;
  for (long long n : {1024LL, 16384LL, 131072LL, 262144LL}) {   // <= 1024 CTAs: all resident (the kernel spins)
    const int reps = 200;
    for (int w = 0; w < 5; w++) {
      epoch++;
      pushExchange<<<(unsigned)((n + 255) / 256), 256>>>(src, peerData, n, done, myFlag2, peerFlag2, epoch);
    }
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int r = 0; r < reps; r++) {
      epoch++;
      pushExchange<<<(unsigned)((n + 255) / 256), 256>>>(src, peerData, n, done, myFlag2, peerFlag2, epoch);
    }
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, a, b));
    if (me == 0)
      printf("ipc_probe: push of %8lld doubles + flag + wait: %.2f us per exchange (%.1f GB/s per direction)\n", n,
             ms * 1e3 / reps, n * 8.0 / (ms * 1e-3 / reps) / 1e9);
  }
  { int to = 0; CK(cudaMemcpyFromSymbol(&to, g_timeouts, 4)); fprintf(stderr, "[%d] total timeouts=%d\n", me, to); }
  // (3) push-only kernels (rank 0 pushes, rank 1 idles): store width / stores in flight / fence, and the copy engine
  if (write(wr, &c, 1) != 1 || read(rd, &c, 1) != 1) return 1;
  if (me == 0) {
    for (long long n : {131072LL, 262144LL, 1048576LL, 4194304LL}) {
      const int reps = 100;
      auto timeIt = [&](const char* name, auto launch) {
        for (int w = 0; w < 3; w++) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        for (int r = 0; r < reps; r++) launch();
        CK(cudaEventRecord(b));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, a, b));
        printf("ipc_probe: %-22s %8lld doubles: %7.2f us  %6.1f GB/s\n", name, n, ms * 1e3 / reps, n * 8.0 / (ms * 1e-3 / reps) / 1e9);
      };
      timeIt("8 B/thread + fence", [&] { pushV1<<<(unsigned)((n + 255) / 256), 256>>>(src, peerData, n); });
      timeIt("8 B/thread no fence", [&] { pushNoFence<<<(unsigned)((n + 255) / 256), 256>>>(src, peerData, n); });
      timeIt("16 B/thread + fence", [&] { pushV2<<<(unsigned)((n / 2 + 255) / 256), 256>>>((const double2*)src, (double2*)peerData, n / 2); });
      timeIt("4x16 B/thread + fence", [&] { pushV2x4<<<(unsigned)((n / 8 + 255) / 256), 256>>>((const double2*)src, (double2*)peerData, n / 2); });
      timeIt("8 B/thread, local dst", [&] { pushV1<<<(unsigned)((n + 255) / 256), 256>>>(src, myData, n); });
      timeIt("cudaMemcpyAsync D2D", [&] { cudaMemcpyAsync(peerData, src, n * 8, cudaMemcpyDeviceToDevice, 0); });
    }
  }
  if (write(wr, &c, 1) != 1 || read(rd, &c, 1) != 1) return 1;
  // check the data arrived
  double hv = 0;
  CK(cudaMemcpy(&hv, myData + 5, 8, cudaMemcpyDeviceToHost));
  double expect;
  memset(&expect, 1, 8);
  if (me == 0) printf("ipc_probe: received data %s\n", hv == expect ? "OK" : "MISMATCH");
  if (write(wr, &c, 1) != 1 || read(rd, &c, 1) != 1) return 1;
  CK(cudaIpcCloseMemHandle(peer));
  if (me == 0) { int st; waitpid(pid, &st, 0); }
  return 0;
}
