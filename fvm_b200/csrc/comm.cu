// fvm_b200 / libfvmgpu -- halo exchange + reductions between ranks (see comm.cuh).
#include "comm.cuh"
#include "../../include/fvmgpu.h"

#ifndef FVMGPU_HOSTSIM
#include <dlfcn.h>
#endif

namespace fvmgpu {

struct HaloPackKernel {  // send[k*w + c] = x[idx[k]*w + c]
  const int* idx; const double* x; double* send; int w;
  FVM_DEV void operator()(long long t) const {
    const long long k = t / w; const int c = (int)(t - k * w);
    send[t] = x[(long long)idx[k] * w + c];
  }
};
struct HaloUnpackKernel {
  const int* idx; const double* recv; double* x; int w;
  FVM_DEV void operator()(long long t) const {
    const long long k = t / w; const int c = (int)(t - k * w);
    x[(long long)idx[k] * w + c] = recv[t];
  }
};

void Halo::detectContiguous(const std::vector<int>& gather) {
  gatherBase = -1;
  if (gather.empty()) return;
  int expect = 0;
  for (const HaloMsg& hm : msgs) {  // message order must also be slot order
    if (hm.recvOff != expect) return;
    expect += hm.recvCnt;
  }
  for (size_t k = 1; k < gather.size(); k++)
    if (gather[k] != gather[0] + (int)k) return;
  gatherBase = gather[0];
}

void Halo::build(const std::vector<HaloMsg>& m, const std::vector<int>& scatter, const std::vector<int>& gather) {
  msgs = m;
  nSend = (int)scatter.size();
  nRecv = (int)gather.size();
  scatterIdx.upload(scatter.data(), scatter.size());
  gatherIdx.upload(gather.data(), gather.size());
  // staging buffers for the common width-1 exchange are allocated here, not lazily: the first
  // exchange may happen inside a CUDA-graph capture, where an allocation would become a graph node
  sendBuf.alloc((size_t)nSend + 1);
  recvBuf.alloc((size_t)nRecv + 1);
  widthCap = 1;
  detectContiguous(gather);
  peerPlanBuild(peer, msgs, 1);   // NVLink peer stores when the transport is up (else the plan stays invalid: NCCL)
}

void Halo::buildDev(const std::vector<HaloMsg>& m, DBuf<int>&& scatterDev, int nSendEntries, const std::vector<int>& gather) {
  msgs = m;
  nSend = nSendEntries;
  nRecv = (int)gather.size();
  scatterIdx = std::move(scatterDev);
  gatherIdx.upload(gather.data(), gather.size());
  sendBuf.alloc((size_t)nSend + 1);
  recvBuf.alloc((size_t)nRecv + 1);
  widthCap = 1;
  detectContiguous(gather);
  peerPlanBuild(peer, msgs, 1);
}

struct HaloIotaKernel { int base; int* p; FVM_DEV void operator()(long long i) const { p[i] = base + (int)i; } };
void Halo::buildDevContiguous(const std::vector<HaloMsg>& m, DBuf<int>&& scatterDev, int nSendEntries, int base, int nRecvEntries) {
  msgs = m;
  nSend = nSendEntries;
  nRecv = nRecvEntries;
  scatterIdx = std::move(scatterDev);
  gatherIdx.alloc((size_t)nRecv + 1);
  if (nRecv) parallelFor(nRecv, HaloIotaKernel{base, gatherIdx.p});
  sendBuf.alloc((size_t)nSend + 1);
  recvBuf.alloc((size_t)nRecv + 1);
  widthCap = 1;
  gatherBase = -1;
  int expect = 0;
  bool inOrder = true;
  for (const HaloMsg& hm : msgs) { if (hm.recvOff != expect) inOrder = false; expect += hm.recvCnt; }
  if (inOrder && nRecv > 0) gatherBase = base;
  peerPlanBuild(peer, msgs, 1);
}

void Halo::exchange(double* x, int width) {
  if (!commActive() || msgs.empty()) return;
  if (peer.valid()) {
    // gather + store into the neighbours' memory + flag + wait + unpack: ONE kernel, no host-side collective call
    if (width > peer.width && !peerPlanBuild(peer, msgs, width)) fail("halo exchange: no room in the peer windows for width %d (FVMGPU_PEER_WINDOW_MB)", width);
    peerExchange(peer, scatterIdx.p, gatherIdx.p, gatherBase, x, x, width);
    return;
  }
  if (width > widthCap) {
    sendBuf.alloc((size_t)nSend * width + 1);
    recvBuf.alloc((size_t)nRecv * width + 1);
    widthCap = width;
  }
  if (nSend) parallelFor((long long)nSend * width, HaloPackKernel{scatterIdx.p, x, sendBuf.p, width});
  if (gatherBase >= 0) {  // ghost slots are one contiguous run in message order: receive straight into x
    commExchange(msgs, sendBuf.p, x + (size_t)gatherBase * width, width);
    return;
  }
  commExchange(msgs, sendBuf.p, recvBuf.p, width);
  if (nRecv) parallelFor((long long)nRecv * width, HaloUnpackKernel{gatherIdx.p, recvBuf.p, x, width});
}

void Halo::exchangeBegin(double* x) {
  if (!commActive() || msgs.empty()) return;
  peerExchangeBegin(peer, scatterIdx.p, x, 1);
}
void Halo::exchangeEnd(double* x) {
  if (!commActive() || msgs.empty()) return;
  peerExchangeEnd(peer, gatherIdx.p, gatherBase, x, 1);
}

#ifndef FVMGPU_HOSTSIM
// ---------------------------------------------------------------- NCCL (dlopen)
namespace {
struct Id128 { char b[128]; };  // ncclUniqueId is passed BY VALUE (128 bytes)
enum { kNcclChar = 0, kNcclDouble = 8, kNcclSum = 0 };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
};
NcclApi& nccl() {
  static NcclApi api;
  if (!api.lib) {
    // in a torchrun-launched process torch has already loaded its bundled libnccl.so.2
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) fail("multi-GPU: cannot dlopen libnccl.so.2 (%s)", dlerror());
#define NCCL_SYM(field, name) api.field = (decltype(api.field))dlsym(api.lib, name); \
    if (!api.field) fail("multi-GPU: libnccl lacks %s", name)
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NCCL_SYM(GroupStart, "ncclGroupStart");
    NCCL_SYM(GroupEnd, "ncclGroupEnd");
    NCCL_SYM(Send, "ncclSend");
    NCCL_SYM(Recv, "ncclRecv");
    NCCL_SYM(AllReduce, "ncclAllReduce");
    NCCL_SYM(AllGather, "ncclAllGather");
#undef NCCL_SYM
    api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
  }
  return api;
}
void ncclCheck(int rc, const char* what) {
  if (rc) fail("NCCL %s failed: %s", what, nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
}
}  // namespace

void commUniqueId(void* out128) { ncclCheck(nccl().GetUniqueId(out128), "ncclGetUniqueId"); }
void commInitNccl(int nranks, int rank, const void* uniqueId128) {
  if (nranks > 1) {
    Id128 id;
    std::memcpy(id.b, uniqueId128, 128);
    void* comm = nullptr;
    ncclCheck(nccl().CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
    ctx().ncclComm = comm;
  }
}
void commDestroy() {
  peerShutdown();
  if (ctx().ncclComm) { nccl().CommDestroy(ctx().ncclComm); ctx().ncclComm = nullptr; }
}

void commExchange(const std::vector<HaloMsg>& msgs, const double* send_d, double* recv_d, int width) {
  if (!commActive()) return;
  NcclApi& n = nccl();
  void* comm = ctx().ncclComm;
  if (!comm) fail("multi-GPU: communicator not initialised (fvmgpu_comm_init)");
  ncclCheck(n.GroupStart(), "ncclGroupStart");
  for (const HaloMsg& m : msgs) {
    if (m.sendCnt) ncclCheck(n.Send(send_d + (size_t)m.sendOff * width, (size_t)m.sendCnt * width, kNcclDouble, m.rank,
                                    comm, ctx().stream), "ncclSend");
    if (m.recvCnt) ncclCheck(n.Recv(recv_d + (size_t)m.recvOff * width, (size_t)m.recvCnt * width, kNcclDouble, m.rank,
                                    comm, ctx().stream), "ncclRecv");
  }
  ncclCheck(n.GroupEnd(), "ncclGroupEnd");
  ctx().collectives++;
}
void commAllreduceSum(double* data_d, int cnt) {
  if (!commActive()) return;
  if (peerAllreduceSum(data_d, cnt)) return;
  ncclCheck(nccl().AllReduce(data_d, data_d, (size_t)cnt, kNcclDouble, kNcclSum, ctx().ncclComm, ctx().stream),
            "ncclAllReduce");
  ctx().collectives++;
}
void commAllgather(const void* send_d, void* recv_d, size_t bytesPerRank) {
  if (!commActive()) { copyD2D(recv_d, send_d, bytesPerRank); return; }
  ncclCheck(nccl().AllGather(send_d, recv_d, bytesPerRank, kNcclChar, ctx().ncclComm, ctx().stream), "ncclAllGather");
  ctx().collectives++;
}
#else
// ---------------------------------------------------------------- hostsim transport (tests only)
namespace {
fvmgpu_hostsim_exchange_fn g_exchange = nullptr;
fvmgpu_hostsim_allreduce_fn g_allreduce = nullptr;
fvmgpu_hostsim_allgather_fn g_allgather = nullptr;
}
void hostsimSetComm(fvmgpu_hostsim_exchange_fn e, fvmgpu_hostsim_allreduce_fn r, fvmgpu_hostsim_allgather_fn g) {
  g_exchange = e; g_allreduce = r; g_allgather = g;
}
void commUniqueId(void* out128) { std::memset(out128, 0, 128); }
void commInitNccl(int, int, const void*) {}
void commDestroy() {}
void commExchange(const std::vector<HaloMsg>& msgs, const double* send_d, double* recv_d, int width) {
  if (!commActive()) return;
  if (!g_exchange) fail("hostsim: no exchange callback registered");
  std::vector<int> peer, so, sc, ro, rc;
  for (const HaloMsg& m : msgs) {
    peer.push_back(m.rank);
    so.push_back(m.sendOff * width); sc.push_back(m.sendCnt * width);
    ro.push_back(m.recvOff * width); rc.push_back(m.recvCnt * width);
  }
  g_exchange((int)msgs.size(), peer.data(), so.data(), sc.data(), send_d, ro.data(), rc.data(), recv_d);
  ctx().collectives++;
}
void commAllreduceSum(double* data_d, int cnt) {
  if (!commActive()) return;
  if (!g_allreduce) fail("hostsim: no allreduce callback registered");
  g_allreduce(data_d, cnt);
  ctx().collectives++;
}
void commAllgather(const void* send_d, void* recv_d, size_t bytesPerRank) {
  if (!commActive()) { copyD2D(recv_d, send_d, bytesPerRank); return; }
  if (!g_allgather) fail("hostsim: no allgather callback registered");
  g_allgather(send_d, recv_d, (long long)bytesPerRank);
  ctx().collectives++;
}
#endif

std::vector<double> commGatherHost(const double* vals, int cnt) {
  const int nr = ctx().nranks;
  std::vector<double> out((size_t)nr * cnt);
  if (!commActive()) { for (int k = 0; k < cnt; k++) out[(size_t)k] = vals[k]; return out; }
  DBuf<double> d((size_t)cnt), all((size_t)nr * cnt);
  copyH2D(d.p, vals, sizeof(double) * (size_t)cnt);
  commAllgather(d.p, all.p, sizeof(double) * (size_t)cnt);
  copyD2H(out.data(), all.p, sizeof(double) * out.size());
  return out;
}

double commSumHost(double v) {
  if (!commActive()) return v;
  DBuf<double> d(1);
  copyH2D(d.p, &v, sizeof(double));
  commAllreduceSum(d.p, 1);
  double out;
  copyD2H(&out, d.p, sizeof(double));
  return out;
}

double commMaxHost(double v) {
  if (!commActive()) return v;
  const int nr = ctx().nranks;
  DBuf<double> d(1), all((size_t)nr);
  copyH2D(d.p, &v, sizeof(double));
  commAllgather(d.p, all.p, sizeof(double));
  std::vector<double> h = all.toHost();
  double m = h[0];
  for (double x : h) m = x > m ? x : m;
  return m;
}

}  // namespace fvmgpu
