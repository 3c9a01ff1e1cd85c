// fvm_b200 / libfvmgpu -- NVLink peer-memory transport (protocol and layout: peer.cuh).
#include "comm.cuh"

#include <map>

namespace fvmgpu {

PeerPlan& PeerPlan::operator=(PeerPlan&& o) noexcept {
  if (this != &o) {
    release();
    msgs = std::move(o.msgs); counters = std::move(o.counters);
    nMsgs = o.nMsgs; width = o.width; maxSend = o.maxSend; maxRecv = o.maxRecv; generation = o.generation;
    totalSend = o.totalSend; totalRecv = o.totalRecv;
    slices = std::move(o.slices);
    o.nMsgs = 0; o.width = 0; o.slices.clear();
  }
  return *this;
}

#ifdef FVMGPU_HOSTSIM
// the test-only host simulator has no peer memory: its transport is the callback set of comm.cu
void PeerPlan::release() { slices.clear(); width = 0; nMsgs = 0; }
bool PeerPlan::valid() const { return false; }
bool peerActive() { return false; }
void peerInit() {}
void peerShutdown() {}
void peerCheck() {}
bool peerPlanBuild(PeerPlan&, const std::vector<HaloMsg>&, int) { return false; }
void peerExchange(PeerPlan&, const int*, const int*, int, const double*, double*, int) { fail("peer transport: device build only"); }
void peerExchangeBegin(PeerPlan&, const int*, const double*, int) { fail("peer transport: device build only"); }
void peerExchangeEnd(PeerPlan&, const int*, int, double*, int) { fail("peer transport: device build only"); }
bool peerAllreduceSum(double*, int) { return false; }
bool peerGatherPlan(PeerPlan&, long long) { return false; }
void peerAllgather(PeerPlan&, const double*, double*, long long) { fail("peer transport: device build only"); }
#else

namespace {
constexpr size_t kControlBytes = 64 << 10;
constexpr int kReduceMax = 8;  // doubles per all-reduce

struct WindowAlloc {           // first-fit allocator over my window inside ONE peer's arena (host side, sender owned)
  std::map<long long, size_t> freeBlocks;   // offset -> bytes
  void reset(size_t bytes) { freeBlocks.clear(); freeBlocks[0] = bytes; }
  long long alloc(size_t bytes) {
    bytes = (bytes + kPeerGranule - 1) / kPeerGranule * kPeerGranule;
    for (auto it = freeBlocks.begin(); it != freeBlocks.end(); ++it) {
      if (it->second < bytes) continue;
      const long long off = it->first;
      const size_t rest = it->second - bytes;
      freeBlocks.erase(it);
      if (rest) freeBlocks[off + (long long)bytes] = rest;
      return off;
    }
    return -1;
  }
  void free(long long off, size_t bytes) {
    bytes = (bytes + kPeerGranule - 1) / kPeerGranule * kPeerGranule;
    auto it = freeBlocks.emplace(off, bytes).first;
    auto nx = std::next(it);
    if (nx != freeBlocks.end() && it->first + (long long)it->second == nx->first) { it->second += nx->second; freeBlocks.erase(nx); }
    if (it != freeBlocks.begin()) {
      auto pv = std::prev(it);
      if (pv->first + (long long)pv->second == it->first) { pv->second += it->second; freeBlocks.erase(it); }
    }
  }
};

struct PeerState {
  bool active = false;
  unsigned generation = 0;               // bumped by every peerInit: plans of an earlier arena are dead
  int nranks = 1, rank = 0;
  size_t windowBytes = 0, arenaBytes = 0;
  char* arena = nullptr;                 // mine
  std::vector<char*> peerArena;          // [rank] -> mapped base (mine at [rank])
  std::vector<WindowAlloc> out;          // [dst]: my window inside dst's arena
  unsigned long long* seq = nullptr;     // device: [nranks] messages exchanged per peer
  int* err = nullptr;                    // device: sticky timeout flag
  PeerPlan reducePlan;
};
PeerState& ps() { static PeerState s; return s; }

char* windowBase(int arenaOf, int writer) { return ps().peerArena[(size_t)arenaOf] + kControlBytes + (size_t)writer * ps().windowBytes; }
unsigned long long* flagOf(int arenaOf, int writer) {
  return reinterpret_cast<unsigned long long*>(ps().peerArena[(size_t)arenaOf]) + (size_t)writer * kPeerFlagStride;
}
}  // namespace

bool peerActive() { return ps().active; }

bool PeerPlan::valid() const { return width > 0 && ps().active && generation == ps().generation; }

void PeerPlan::release() {
  if (ps().active && generation == ps().generation)
    for (const Slice& s : slices) ps().out[(size_t)s.rank].free(s.off, s.bytes);
  slices.clear();
  msgs.release(); counters.release();
  width = 0; nMsgs = 0;
}

void peerShutdown() {
  PeerState& S = ps();
  if (!S.arena) return;
  cudaStreamSynchronize(ctx().stream);
  S.reducePlan.release();
  S.active = false;
  for (int r = 0; r < S.nranks; r++)
    if (r != S.rank && S.peerArena[(size_t)r]) cudaIpcCloseMemHandle(S.peerArena[(size_t)r]);
  S.peerArena.clear();
  // an exported allocation must outlive every mapping of it: all ranks unmap, agree, then free (collective, the
  // NCCL communicator is still alive here -- commDestroy calls this first)
  if (ctx().ncclComm) (void)commSumHost(1.0);
  cudaFree(S.arena); S.arena = nullptr;
  if (S.seq) { cudaFree(S.seq); S.seq = nullptr; }
  if (S.err) { cudaFree(S.err); S.err = nullptr; }
}

void peerInit() {
  PeerState& S = ps();
  peerShutdown();
  S.nranks = ctx().nranks; S.rank = ctx().rank;
  S.generation++;
  if (S.nranks < 2 || S.nranks > kPeerMaxMsgs) return;
  if (const char* e = getenv("FVMGPU_PEER")) { if (atoi(e) == 0) return; }   // A/B switch: 0 = stay on NCCL
  double mb = 64;
  if (const char* e = getenv("FVMGPU_PEER_WINDOW_MB")) mb = atof(e);
  if (mb < 1) mb = 1;
  if (mb > 2048) mb = 2048;   // flag payload: 24 bits of 256-byte granules
  S.windowBytes = ((size_t)(mb * 1048576.0) + 4095) / 4096 * 4096;
  S.arenaBytes = kControlBytes + S.windowBytes * (size_t)S.nranks;
  // 1 = this rank can offer and map peer memory; agreed by all-reduce so that every rank takes the same path
  int ok = 1;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (cudaMalloc((void**)&S.arena, S.arenaBytes) != cudaSuccess) { (void)cudaGetLastError(); S.arena = nullptr; ok = 0; }
  if (ok && cudaMemset(S.arena, 0, kControlBytes) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, S.arena) != cudaSuccess) { (void)cudaGetLastError(); ok = 0; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  const size_t rec = 64 + 8;  // handle + ok flag + device ordinal
  DBuf<char> sendD(rec), recvD(rec * S.nranks);
  std::vector<char> sendH(rec, 0), recvH(rec * S.nranks);
  std::memcpy(sendH.data(), &mine, 64);
  int meta[2] = {ok, ctx().device};
  std::memcpy(sendH.data() + 64, meta, 8);
  copyH2D(sendD.p, sendH.data(), rec);
  commAllgather(sendD.p, recvD.p, rec);
  copyD2H(recvH.data(), recvD.p, rec * S.nranks);
  for (int r = 0; r < S.nranks; r++) {
    int m[2];
    std::memcpy(m, recvH.data() + (size_t)r * rec + 64, 8);
    if (!m[0]) ok = 0;
    if (r != S.rank) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, ctx().device, m[1]) != cudaSuccess || !can) { (void)cudaGetLastError(); ok = 0; }
    }
  }
  S.peerArena.assign((size_t)S.nranks, nullptr);
  if (ok) {
    for (int r = 0; r < S.nranks && ok; r++) {
      if (r == S.rank) { S.peerArena[(size_t)r] = S.arena; continue; }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, recvH.data() + (size_t)r * rec, 64);
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); ok = 0; }
      S.peerArena[(size_t)r] = (char*)p;
    }
  }
  // every rank must have mapped every arena before anybody stores into one: agree (also serves as the barrier)
  ok = commSumHost(ok ? 1.0 : 0.0) > S.nranks - 0.5 ? 1 : 0;
  if (!ok) {
    for (int r = 0; r < S.nranks; r++)
      if (r != S.rank && S.peerArena[(size_t)r]) cudaIpcCloseMemHandle(S.peerArena[(size_t)r]);
    S.peerArena.clear();
    if (S.arena) { cudaFree(S.arena); S.arena = nullptr; }
    return;   // NCCL transport stays in charge
  }
  CUDA_CHECK(cudaMalloc((void**)&S.seq, sizeof(unsigned long long) * (size_t)S.nranks));
  CUDA_CHECK(cudaMemset(S.seq, 0, sizeof(unsigned long long) * (size_t)S.nranks));
  CUDA_CHECK(cudaMalloc((void**)&S.err, sizeof(int)));
  CUDA_CHECK(cudaMemset(S.err, 0, sizeof(int)));
  S.out.assign((size_t)S.nranks, WindowAlloc());
  for (auto& w : S.out) w.reset(S.windowBytes);
  S.active = true;
  std::vector<HaloMsg> all;
  for (int r = 0; r < S.nranks; r++) {
    if (r == S.rank) continue;
    HaloMsg m; m.rank = r; m.sendOff = 0; m.sendCnt = kReduceMax; m.recvOff = 0; m.recvCnt = kReduceMax;
    all.push_back(m);
  }
  if (!peerPlanBuild(S.reducePlan, all, 1)) { S.active = false; return; }
}

void peerCheck() {
  PeerState& S = ps();
  if (!S.active) return;
  int e = 0;
  CUDA_CHECK(cudaMemcpyAsync(&e, S.err, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
  CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
  if (e) fail("multi-GPU: a rank waited more than a minute for a neighbour's halo message (peer transport timeout)");
}

bool peerPlanBuild(PeerPlan& P, const std::vector<HaloMsg>& msgs, int width) {
  PeerState& S = ps();
  P.release();
  if (!S.active || width < 1) return false;
  if (msgs.size() > (size_t)kPeerMaxMsgs) return false;
  P.generation = S.generation;
  std::vector<PeerMsg> h;
  long long maxSend = 0, maxRecv = 0;
  for (const HaloMsg& m : msgs) {
    if (m.rank < 0 || m.rank >= S.nranks || m.rank == S.rank) fail("peer plan: bad neighbour rank %d", m.rank);
    const size_t bytes = (size_t)(m.sendCnt > 0 ? m.sendCnt : 1) * width * sizeof(double);
    PeerMsg d;
    d.sendOff = m.sendOff; d.sendCnt = m.sendCnt; d.recvOff = m.recvOff; d.recvCnt = m.recvCnt;
    for (int c = 0; c < 2; c++) {
      const long long off = S.out[(size_t)m.rank].alloc(bytes);
      if (off < 0) { P.release(); return false; }   // window full: this pattern stays on NCCL
      P.slices.push_back(PeerPlan::Slice{m.rank, off, bytes});
      d.remote[c] = reinterpret_cast<double*>(windowBase(m.rank, S.rank) + off);
      d.remoteOff[c] = (unsigned long long)(off / kPeerGranule);
    }
    d.remoteFlag = flagOf(m.rank, S.rank);
    d.localFlag = flagOf(S.rank, m.rank);
    d.localWindow = windowBase(S.rank, m.rank);
    d.seq = S.seq + m.rank;
    h.push_back(d);
    maxSend = std::max<long long>(maxSend, m.sendCnt);
    maxRecv = std::max<long long>(maxRecv, m.recvCnt);
  }
  P.nMsgs = (int)h.size();
  P.width = width;
  P.maxSend = maxSend; P.maxRecv = maxRecv;
  P.totalSend = P.totalRecv = 0;
  for (const HaloMsg& m : msgs) { P.totalSend += m.sendCnt; P.totalRecv += m.recvCnt; }
  P.msgs.upload(h.data(), h.size());
  P.counters.alloc(2);
  P.counters.zero();
  return true;
}

// ---------------------------------------------------------------- kernels
// One kernel per exchange; the grid is at most one CTA per SM, so every CTA is resident and may spin.
//   1. push   : my scatter entries -> the neighbours' windows (coalesced stores over NVLink)
//   2. signal : the CTA that finishes last stores one flag per neighbour (sequence number + where the data is)
//   3. wait   : every CTA spins on the neighbours' flags in OUR control block (local memory)
//   4. unpack : the neighbours' values, read from our window with L2-only loads, -> the ghost slots
//   5. the CTA that finishes last advances the per-neighbour sequence numbers for the next kernel
__global__ void __launch_bounds__(256) k_peer_exchange(const PeerMsg* msgs, int nMsgs, const int* scatterIdx,
                                                        const int* gatherIdx, int gatherBase, const double* xSend,
                                                        double* xRecv, int width, unsigned* counters, int* err) {
  __shared__ unsigned long long sOff[kPeerMaxMsgs];
  __shared__ unsigned long long sSeq[kPeerMaxMsgs];
  __shared__ int sLast;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (threadIdx.x < nMsgs) sSeq[threadIdx.x] = *msgs[threadIdx.x].seq + 1ULL;
  __syncthreads();
  for (int q = 0; q < nMsgs; q++) {
    const PeerMsg m = msgs[q];
    double* dst = m.remote[sSeq[q] & 1ULL];
    const long long total = (long long)m.sendCnt * width;
    if (width == 1) {
      for (long long t = tid; t < total; t += stride) dst[t] = xSend[scatterIdx ? scatterIdx[m.sendOff + t] : m.sendOff + t];
    } else {
      for (long long t = tid; t < total; t += stride) {
        const long long e = t / width;
        const int c = (int)(t - e * width);
        dst[t] = xSend[(long long)(scatterIdx ? scatterIdx[m.sendOff + e] : m.sendOff + e) * width + c];
      }
    }
  }
  // ONE system-scope fence per CTA, by the thread that then counts the CTA in: the barrier orders every thread's
  // stores before it (a fence in every thread throttles the NVLink stores to a third: tools/ipc_probe.cu)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    sLast = atomicAdd(&counters[0], 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (sLast) {
    __threadfence_system();
    if (threadIdx.x < nMsgs) peerSignal(msgs[threadIdx.x], sSeq[threadIdx.x]);
  }
  if (threadIdx.x < nMsgs) sOff[threadIdx.x] = peerWait(msgs[threadIdx.x], sSeq[threadIdx.x], err);
  __syncthreads();
  for (int q = 0; q < nMsgs; q++) {
    const PeerMsg m = msgs[q];
    const double* src = reinterpret_cast<const double*>(m.localWindow + sOff[q]);
    const long long total = (long long)m.recvCnt * width;
    if (width == 1) {
      for (long long t = tid; t < total; t += stride)
        xRecv[gatherBase >= 0 ? gatherBase + m.recvOff + t : gatherIdx[m.recvOff + t]] = __ldcg(src + t);
    } else {
      for (long long t = tid; t < total; t += stride) {
        const long long e = t / width;
        const int c = (int)(t - e * width);
        const long long slot = gatherBase >= 0 ? gatherBase + m.recvOff + e : gatherIdx[m.recvOff + e];
        xRecv[slot * width + c] = __ldcg(src + t);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) sLast = atomicAdd(&counters[1], 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (sLast) {
    if (threadIdx.x < nMsgs) *msgs[threadIdx.x].seq = sSeq[threadIdx.x];
    if (threadIdx.x == 0) { counters[0] = 0; counters[1] = 0; }
  }
}

// the two halves of k_peer_exchange as kernels of their own (see peerExchangeBegin / End)
__global__ void __launch_bounds__(256) k_peer_push(const PeerMsg* msgs, int nMsgs, const int* scatterIdx, const double* xSend,
                                                    int width, unsigned* counters) {
  __shared__ unsigned long long sSeq[kPeerMaxMsgs];
  __shared__ int sLast;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (threadIdx.x < nMsgs) sSeq[threadIdx.x] = *msgs[threadIdx.x].seq + 1ULL;
  __syncthreads();
  // every gather is a chain of dependent loads (index -> value) ending in a store that crosses NVLink: the grid is
  // sized for ONE entry per thread so that all chains are in flight together (this kernel never waits, any grid
  // size is safe); the loop only matters when the caller caps the grid
  for (int q = 0; q < nMsgs; q++) {
    const PeerMsg m = msgs[q];
    double* dst = m.remote[sSeq[q] & 1ULL];
    if (width == 1) {
      for (long long t = tid; t < m.sendCnt; t += stride) dst[t] = xSend[scatterIdx ? scatterIdx[m.sendOff + t] : m.sendOff + t];
    } else {
      const long long total = (long long)m.sendCnt * width;
      for (long long t = tid; t < total; t += stride) {
        const int e = (int)(t / width), c = (int)(t - (long long)e * width);
        dst[t] = xSend[(long long)(scatterIdx ? scatterIdx[m.sendOff + e] : m.sendOff + e) * width + c];
      }
    }
  }
  // ONE system-scope fence per CTA, by the thread that then counts the CTA in: the barrier orders every thread's
  // stores before it (a fence in every thread throttles the NVLink stores to a third: tools/ipc_probe.cu)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    sLast = atomicAdd(&counters[0], 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (sLast) {
    __threadfence_system();
    if (threadIdx.x < nMsgs) peerSignal(msgs[threadIdx.x], sSeq[threadIdx.x]);
    if (threadIdx.x == 0) counters[0] = 0;
  }
}
__global__ void __launch_bounds__(256) k_peer_wait_unpack(const PeerMsg* msgs, int nMsgs, const int* gatherIdx, int gatherBase,
                                                           double* xRecv, int width, unsigned* counters, int* err) {
  __shared__ unsigned long long sOff[kPeerMaxMsgs];
  __shared__ unsigned long long sSeq[kPeerMaxMsgs];
  __shared__ int sLast;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (threadIdx.x < nMsgs) {
    sSeq[threadIdx.x] = *msgs[threadIdx.x].seq + 1ULL;
    sOff[threadIdx.x] = peerWait(msgs[threadIdx.x], sSeq[threadIdx.x], err);
  }
  __syncthreads();
  // (a CTA waits for the NEIGHBOURS' push kernels only, never for a CTA of this grid: any grid size is safe)
  for (int q = 0; q < nMsgs; q++) {
    const PeerMsg m = msgs[q];
    const double* src = reinterpret_cast<const double*>(m.localWindow + sOff[q]);
    if (width == 1) {
      for (long long t = tid; t < m.recvCnt; t += stride)
        xRecv[gatherBase >= 0 ? gatherBase + m.recvOff + t : gatherIdx[m.recvOff + t]] = __ldcg(src + t);
    } else {
      const long long total = (long long)m.recvCnt * width;
      for (long long t = tid; t < total; t += stride) {
        const int e = (int)(t / width), c = (int)(t - (long long)e * width);
        const long long slot = gatherBase >= 0 ? gatherBase + m.recvOff + e : gatherIdx[m.recvOff + e];
        xRecv[slot * width + c] = __ldcg(src + t);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) sLast = atomicAdd(&counters[1], 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (sLast) {
    if (threadIdx.x < nMsgs) *msgs[threadIdx.x].seq = sSeq[threadIdx.x];
    if (threadIdx.x == 0) counters[1] = 0;
  }
}

// sum over the ranks of up to kReduceMax doubles, added in RANK order on every rank: all ranks get the same bits
__global__ void __launch_bounds__(64) k_peer_allreduce(const PeerMsg* msgs, int nMsgs, int myRank, double* data, int n,
                                                        int* err) {
  __shared__ double sVal[kPeerMaxMsgs + 1][kReduceMax];   // [rank][i] (nMsgs + 1 ranks)
  const int t = threadIdx.x;
  unsigned long long k = 0;
  if (t < nMsgs) {
    const PeerMsg m = msgs[t];
    k = *m.seq + 1ULL;
    double* dst = m.remote[k & 1ULL];
    for (int i = 0; i < n; i++) dst[i] = data[i];
    __threadfence_system();
    peerSignal(m, k);
    const unsigned long long off = peerWait(m, k, err);
    const double* src = reinterpret_cast<const double*>(m.localWindow + off);
    // neighbour list of the reduce plan is every other rank in ascending order: message t is rank t (+1 past me)
    const int r = t < myRank ? t : t + 1;
    for (int i = 0; i < n; i++) sVal[r][i] = __ldcg(src + i);
    *m.seq = k;
  }
  if (t < n) sVal[myRank][t] = data[t];
  __syncthreads();
  if (t < n) {
    double s = 0.0;
    for (int r = 0; r <= nMsgs; r++) s += sVal[r][t];
    data[t] = s;
  }
}

// One entry per thread. The fused kernel spins on the neighbours' flags while its own flags are sent by ITS last
// CTA, so all its CTAs must be resident (capped at what the occupancy calculator says fits); the split kernels
// wait for nothing inside their own grid and are not capped.
static int exchangeGrid(long long work, bool mustBeResident) {
  long long g = (work + 255) / 256;
  if (g < 1) g = 1;
  if (mustBeResident) {
    static int perSm = 0;
    if (!perSm) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_peer_exchange, 256, 0) != cudaSuccess || perSm < 1) perSm = 1;
    }
    const long long cap = (long long)ctx().smCount * perSm;
    if (g > cap) g = cap;
  }
  if (g > 65535LL * 16) g = 65535LL * 16;
  return (int)g;
}

void peerExchange(PeerPlan& P, const int* scatterIdx, const int* gatherIdx, int gatherBase, const double* xSend,
                  double* xRecv, int width) {
  if (!P.valid() || width > P.width) fail("peer exchange: plan not built for width %d", width);
  if (P.nMsgs == 0) return;
  const int grid = exchangeGrid(std::max(P.maxSend, P.maxRecv) * width, true);
  ProfileScope prof("N6fvmgpu15k_peer_exchangeE", std::max(P.maxSend, P.maxRecv) * width);
  k_peer_exchange<<<grid, 256, 0, ctx().stream>>>(P.msgs.p, P.nMsgs, scatterIdx, gatherIdx, gatherBase, xSend, xRecv, width,
                                                  P.counters.p, ps().err);
  ctx().launches++;
  ctx().collectives++;
  CUDA_CHECK(cudaGetLastError());
}

void peerExchangeBegin(PeerPlan& P, const int* scatterIdx, const double* xSend, int width) {
  if (!P.valid() || width > P.width) fail("peer exchange: plan not built for width %d", width);
  if (P.nMsgs == 0) return;
  ProfileScope prof("N6fvmgpu11k_peer_pushE", P.maxSend * width);
  k_peer_push<<<exchangeGrid(P.maxSend * width, false), 256, 0, ctx().stream>>>(P.msgs.p, P.nMsgs, scatterIdx, xSend, width, P.counters.p);
  ctx().launches++;
  ctx().collectives++;
  CUDA_CHECK(cudaGetLastError());
}
void peerExchangeEnd(PeerPlan& P, const int* gatherIdx, int gatherBase, double* xRecv, int width) {
  if (!P.valid() || width > P.width) fail("peer exchange: plan not built for width %d", width);
  if (P.nMsgs == 0) return;
  ProfileScope prof("N6fvmgpu18k_peer_wait_unpackE", P.maxRecv * width);
  k_peer_wait_unpack<<<exchangeGrid(P.maxRecv * width, false), 256, 0, ctx().stream>>>(P.msgs.p, P.nMsgs, gatherIdx, gatherBase, xRecv, width,
                                                                                P.counters.p, ps().err);
  ctx().launches++;
  CUDA_CHECK(cudaGetLastError());
}

bool peerAllreduceSum(double* data_d, int n) {
  PeerState& S = ps();
  if (!S.active || n > kReduceMax || n < 1) return false;
  ProfileScope prof("N6fvmgpu16k_peer_allreduceE", n);
  k_peer_allreduce<<<1, 64, 0, ctx().stream>>>(S.reducePlan.msgs.p, S.reducePlan.nMsgs, S.rank, data_d, n, S.err);
  ctx().launches++;
  ctx().collectives++;
  CUDA_CHECK(cudaGetLastError());
  return true;
}

bool peerGatherPlan(PeerPlan& P, long long count) {
  PeerState& S = ps();
  if (!S.active) return false;
  std::vector<HaloMsg> all;
  for (int r = 0; r < S.nranks; r++) {
    if (r == S.rank) continue;
    HaloMsg m; m.rank = r; m.sendOff = 0; m.sendCnt = (int)count; m.recvOff = (int)((long long)r * count); m.recvCnt = (int)count;
    all.push_back(m);
  }
  return peerPlanBuild(P, all, 1);
}

void peerAllgather(PeerPlan& P, const double* send_d, double* recv_d, long long count) {
  // own block first (stream ordered), then the neighbours' blocks land at r * count
  copyD2D(recv_d + (size_t)ps().rank * count, send_d, (size_t)count * sizeof(double));
  peerExchange(P, nullptr, nullptr, 0, send_d, recv_d, 1);
}
#endif

}  // namespace fvmgpu
