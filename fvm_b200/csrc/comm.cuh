// fvm_b200 / libfvmgpu -- inter-rank communication of the hot path (one process per GPU).
//
// What the reference does with MPI                         here
//   MultiField::sync / Field::syncLocal                     Halo::exchange: ONE kernel stores the interface values
//     Isend/Irecv of packed ghost values + Waitall            into the neighbour's memory over NVLink, flags, waits
//     (F/MultiField.cpp:488-551, F/Field.cpp:333-394)         and unpacks (peer.cuh); all stream ordered
//   MultiFieldReduction::reduceSum  Allreduce(SUM)          commAllreduceSum (peer stores, summed in rank order)
//     (F/MultiFieldReduction.cpp:213-225)
//   LinearSystemMerger Gatherv/Scatterv of coarse levels    commAllgather / peerAllgather (equal-size padded blocks)
//     (F/LinearSystemMerger.cpp:720-819)
// Without peer access between the ranks (or with FVMGPU_PEER=0) the same three steps run as pack kernel + grouped
// ncclSend/ncclRecv + unpack kernel, ncclAllReduce and ncclAllGather; NCCL also carries the set-up.
//
// NCCL is resolved with dlopen (no link-time dependency). The FVMGPU_HOSTSIM test build replaces
// the transport by callbacks the test harness registers (tests drive them with torch.distributed
// gloo, world_size 2); the pack/unpack kernels and all index logic are the same code.
#pragma once
#include "common.cuh"
#include "peer.cuh"

namespace fvmgpu {

struct HaloMsg {
  int rank;              // peer
  int sendOff, sendCnt;  // in entries of the send buffer
  int recvOff, recvCnt;
};

inline bool commActive() { return ctx().nranks > 1; }
// exchange `width` doubles per entry with every peer (send_d/recv_d are device buffers)
void commExchange(const std::vector<HaloMsg>& msgs, const double* send_d, double* recv_d, int width);
void commAllreduceSum(double* data_d, int n);                       // in place
void commAllgather(const void* send_d, void* recv_d, size_t bytesPerRank);
double commSumHost(double v);                                       // host scalar convenience (synchronises)
std::vector<double> commGatherHost(const double* vals, int cnt);    // every rank's cnt values, rank-major (one all-gather)
double commMaxHost(double v);
inline bool commAll(bool ok) { return !commActive() ? ok : commSumHost(ok ? 1.0 : 0.0) > ctx().nranks - 0.5; }
inline bool commAny(bool ok) { return !commActive() ? ok : commSumHost(ok ? 1.0 : 0.0) > 0.5; }

// ---- NVLink peer-memory transport (peer.cuh / peer.cu): active after commInitNccl when every rank could map
// every other rank's arena (one node, P2P capable); otherwise everything below stays on NCCL.
struct PeerPlan {   // device descriptors + the slices of the neighbours' windows of ONE exchange pattern
  DBuf<PeerMsg> msgs;
  DBuf<unsigned> counters;             // two arrival counters of the exchange kernel
  int nMsgs = 0;
  int width = 0;                       // doubles per entry the slices were sized for
  unsigned generation = 0;             // arena generation the plan belongs to (a new communicator = a new arena)
  long long maxSend = 0, maxRecv = 0;  // entries of the largest message
  long long totalSend = 0, totalRecv = 0;
  struct Slice { int rank; long long off; size_t bytes; };
  std::vector<Slice> slices;
  PeerPlan() {}
  PeerPlan(const PeerPlan&) = delete;
  PeerPlan& operator=(const PeerPlan&) = delete;
  PeerPlan(PeerPlan&& o) noexcept { *this = std::move(o); }
  PeerPlan& operator=(PeerPlan&& o) noexcept;
  ~PeerPlan() { release(); }
  void release();
  bool valid() const;
};
bool peerActive();
void peerInit();       // collective: arena allocation + exchange of the IPC handles (over the NCCL communicator)
void peerShutdown();
void peerCheck();      // throws when a device-side wait timed out (sticky)
// false (plan left invalid) when a neighbour's window has no room: the caller stays on NCCL
bool peerPlanBuild(PeerPlan& P, const std::vector<HaloMsg>& msgs, int width);
// xRecv[gather slot] <- the neighbours' xSend[scatter entry]; one kernel: push + flag + wait + unpack.
// scatterIdx == nullptr: entries sendOff.. of xSend are sent as they are; gatherBase >= 0: slot = gatherBase + recvOff + k
void peerExchange(PeerPlan& P, const int* scatterIdx, const int* gatherIdx, int gatherBase, const double* xSend,
                  double* xRecv, int width);
// the same in two kernels, so that rows which need no ghost value can run in between: Begin pushes and flags
// (returns without waiting), End waits for the neighbours' flags and unpacks. No other message to the same
// neighbours may be started between the two.
void peerExchangeBegin(PeerPlan& P, const int* scatterIdx, const double* xSend, int width);
void peerExchangeEnd(PeerPlan& P, const int* gatherIdx, int gatherBase, double* xRecv, int width);
bool peerAllreduceSum(double* data_d, int n);   // false: not handled (n too large / transport inactive)
// all-gather of `count` doubles per rank with a plan from peerGatherPlan (recv block r = rank r's send)
bool peerGatherPlan(PeerPlan& P, long long count);
void peerAllgather(PeerPlan& P, const double* send_d, double* recv_d, long long count);

// One halo = scatter/gather index lists + staging buffers for one vector layout.
struct Halo {
  std::vector<HaloMsg> msgs;
  DBuf<int> scatterIdx, gatherIdx;  // entries to send / slots to fill (indices into the vector)
  DBuf<double> sendBuf, recvBuf;
  int nSend = 0, nRecv = 0, widthCap = 0;
  int gatherBase = -1;  // >= 0: the ghost slots are gatherBase .. gatherBase+nRecv-1 in message order -> receive in place
  PeerPlan peer;        // valid: exchanges go over NVLink peer stores instead of NCCL
  bool empty() const { return msgs.empty(); }
  void build(const std::vector<HaloMsg>& m, const std::vector<int>& scatter, const std::vector<int>& gather);
  // same with the scatter list already on the device
  void buildDev(const std::vector<HaloMsg>& m, DBuf<int>&& scatterDev, int nSendEntries, const std::vector<int>& gather);
  // ghost slots gatherBase .. gatherBase + nRecvEntries - 1 in message order (every coarse AMG level): nothing comes from the host
  void buildDevContiguous(const std::vector<HaloMsg>& m, DBuf<int>&& scatterDev, int nSendEntries, int base, int nRecvEntries);
  // x is `width` doubles per entry (AoS); ghost slots of x are overwritten with the peers' values
  void exchange(double* x, int width = 1);
  // split form (peer transport only, see peerExchangeBegin / End)
  bool canSplit() const { return peer.valid(); }
  void exchangeBegin(double* x);
  void exchangeEnd(double* x);
  void detectContiguous(const std::vector<int>& gather);
};

#ifdef FVMGPU_HOSTSIM
// test-only transport: the harness (tests/multirank_worker.py) implements these with torch.distributed (gloo)
extern "C" {
typedef void (*fvmgpu_hostsim_exchange_fn)(int nMsgs, const int* peer, const int* sendOff, const int* sendCnt,
                                           const double* send, const int* recvOff, const int* recvCnt, double* recv);
typedef void (*fvmgpu_hostsim_allreduce_fn)(double* data, int n);
typedef void (*fvmgpu_hostsim_allgather_fn)(const void* send, void* recv, long long bytesPerRank);
}
void hostsimSetComm(fvmgpu_hostsim_exchange_fn e, fvmgpu_hostsim_allreduce_fn r, fvmgpu_hostsim_allgather_fn g);
#endif

void commInitNccl(int nranks, int rank, const void* uniqueId128);
void commUniqueId(void* out128);
void commDestroy();

}  // namespace fvmgpu
