// Probe (measurement tool, not part of the library): does a device-side WHILE loop (CUDA conditional graph node) around
// a V-cycle-like body -- ~60 kernels of ~20 us, one of them a cooperative launch -- beat the host loop
// "launch the cycle graph, read the norm back (blocking D2H), decide"? Prints the time per iteration of both.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cond_graph_probe tools/cond_graph_probe.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)
__global__ void work(double* a, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = a[i] * 0.999 + 1e-3;
}
__global__ void coop(double* a, size_t n) {
  cg::grid_group g = cg::this_grid();
  for (int k = 0; k < 4; k++) {
    for (size_t i = g.thread_rank(); i < n; i += g.size()) a[i] += 1e-6;
    g.sync();
  }
}
__global__ void step(int* it, double* norm) { if (threadIdx.x == 0) { it[0]++; norm[0] *= 0.93; } }
__global__ void test(cudaGraphConditionalHandle h, const int* it, const double* norm, const double* params, double* history) {
  if (threadIdx.x == 0) {
    history[it[0]] = norm[0];
    cudaGraphSetConditional(h, (norm[0] > params[0] && it[0] < (int)params[1]) ? 1u : 0u);
  }
}
int main() {
  cudaStream_t s; CK(cudaStreamCreate(&s));
  const size_t n = 16 << 20;   // 128 MB: ~40 us per pass
  double *a, *norm, *params, *history; int* it;
  CK(cudaMalloc(&a, n * 8)); CK(cudaMemset(a, 0, n * 8));
  CK(cudaMalloc(&norm, 8)); CK(cudaMalloc(&params, 16)); CK(cudaMalloc(&history, 8 * 4096)); CK(cudaMalloc(&it, 4));
  const int K = 60;
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  auto body = [&]() -> int {
    for (int k = 0; k < K; k++) work<<<sms * 8, 256, 0, s>>>(a, n / 8);
    size_t nn = n / 64; void* args[] = {&a, &nn};
    CK(cudaLaunchCooperativeKernel((void*)coop, dim3(sms), dim3(512), args, 0, s));
    step<<<1, 32, 0, s>>>(it, norm);
    return 0;
  };
  auto reset = [&]() -> int {
    int z = 0; double n0 = 1.0, p[2] = {1e-8, 4000};
    CK(cudaMemcpy(it, &z, 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(norm, &n0, 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(params, p, 16, cudaMemcpyHostToDevice));
    return 0;
  };
  // (a) host loop over a captured cycle graph
  cudaGraph_t gc; cudaGraphExec_t gce;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  if (body()) return 1;
  CK(cudaStreamEndCapture(s, &gc));
  CK(cudaGraphInstantiate(&gce, gc, 0));
  for (int rep = 0; rep < 2; rep++) {
    if (reset()) return 1;
    CK(cudaStreamSynchronize(s));
    auto t0 = std::chrono::steady_clock::now();
    int iters = 0; double nv = 1.0;
    while (nv > 1e-8 && iters < 4000) {
      CK(cudaGraphLaunch(gce, s));
      CK(cudaMemcpyAsync(&nv, norm, 8, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      iters++;
    }
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("host loop      : %d iterations, %.3f ms, %.2f us per iteration\n", iters, ms, 1e3 * ms / iters);
  }
  // (b) device WHILE loop
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle h;
  CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {};
  p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node;
  CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
  cudaGraph_t bodyG = p.conditional.phGraph_out[0];
  CK(cudaStreamBeginCaptureToGraph(s, bodyG, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  if (body()) return 1;
  test<<<1, 32, 0, s>>>(h, it, norm, params, history);
  cudaGraph_t out; CK(cudaStreamEndCapture(s, &out));
  cudaGraphExec_t ge; CK(cudaGraphInstantiate(&ge, g, 0));
  for (int rep = 0; rep < 2; rep++) {
    if (reset()) return 1;
    CK(cudaStreamSynchronize(s));
    auto t0 = std::chrono::steady_clock::now();
    CK(cudaGraphLaunch(ge, s));
    CK(cudaStreamSynchronize(s));
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    int iters = 0; double nv = 0;
    CK(cudaMemcpy(&iters, it, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&nv, norm, 8, cudaMemcpyDeviceToHost));
    printf("device WHILE   : %d iterations, %.3f ms, %.2f us per iteration (norm %g)\n", iters, ms, 1e3 * ms / iters, nv);
  }
  return 0;
}
