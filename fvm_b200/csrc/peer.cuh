// fvm_b200 / libfvmgpu -- NVLink peer-memory transport: device-initiated halo exchange, all-gather and
// all-reduce between the ranks of one node WITHOUT a host call or a NCCL kernel on the critical path.
//
// What it replaces (reference: MultiField::sync = packed Isend/Irecv + Waitall, F/MultiField.cpp:488-551;
// MultiFieldMatrix::forwardGS/reverseGS ... x.sync(), F/MultiFieldMatrix.cpp:125-165; reduceSum = Allreduce,
// F/MultiFieldReduction.cpp:213-225): round 1 ran a pack kernel + grouped ncclSend/ncclRecv per exchange
// (~7 + ~15 us of pure latency, ~30 times per V-cycle). Here the values are STORED straight into the
// neighbour's memory over NVLink by the kernel that gathers them, followed by one flag store; the receiver
// spins on its own flag and unpacks. Measured on 2 B200s (tools/ipc_probe.cu): flag one way ~2.7 us,
// push + flag + wait of 128 K doubles in one kernel 13 us.
//
// Memory: every rank cudaMallocs ONE arena at communicator set-up and maps all the others' arenas with
// cudaIpcOpenMemHandle. An arena is  [control block | window written by rank 0 | window by rank 1 | ...].
// The window of rank s inside rank d's arena is written by s ONLY, so s alone decides where inside it a
// message goes (host-side allocator on s, no agreement needed) and tells d with the flag itself:
//     flag value = (sequence number << 24) | (offset inside the window / 256 bytes)
// One flag per ordered pair (s -> d) lives in d's control block; the sequence number counts every message of
// that pair (halo exchanges, all-gathers and all-reduces alike; all of them are bidirectional and issued in
// the same order by both ends, exactly what the NCCL path required as well), is kept in DEVICE memory and
// advanced by the kernels themselves, so that a captured CUDA graph can be replayed.
// Every message buffer is allocated twice and the copy used is sequence & 1: a sender may only run message
// k+1 after it has received the receiver's flag k, which the receiver stores in the kernel that FOLLOWS the
// one that unpacked message k-1 -- so the copy written by k+1 is never still being read (no acknowledgement
// round trip needed).
// A wait gives up after ~60 s and raises a sticky error flag (a dead peer must not hang the GPU for ever);
// the host checks it at the end of every solve.
#pragma once
#include "common.cuh"

namespace fvmgpu {

constexpr int kPeerMaxMsgs = 64;          // messages (= neighbour ranks) per exchange
constexpr int kPeerGranule = 256;         // window offsets are multiples of this many bytes
constexpr int kPeerFlagStride = 16;       // u64 slots between two flags (128 B: one flag per line)

struct PeerMsg {                          // one neighbour of one exchange plan (device)
  int sendOff, sendCnt, recvOff, recvCnt; // in entries, as HaloMsg
  double* remote[2];                      // my slice of the peer's arena for this plan, copy 0 / 1
  unsigned long long remoteOff[2];        // the same as window offsets in granules (sent with the flag)
  unsigned long long* remoteFlag;         // peer's control block: flag[me]
  const unsigned long long* localFlag;    // my control block: flag[peer]
  const char* localWindow;                // my arena: the window the peer writes
  unsigned long long* seq;                // messages exchanged with this peer so far (device counter)
};

#ifndef FVMGPU_HOSTSIM
__device__ __forceinline__ void peerStoreFlag(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peerLoadFlag(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// spin until the peer's message `k` has been flagged; returns the byte offset of its data inside the window
__device__ __forceinline__ unsigned long long peerWait(const PeerMsg& m, unsigned long long k, int* err) {
  unsigned long long v = peerLoadFlag(m.localFlag);
  if ((v >> 24) < k) {
    const long long t0 = clock64();
    for (;;) {
      v = peerLoadFlag(m.localFlag);
      if ((v >> 24) >= k) break;
      if (*(volatile int*)err) break;                                          // sticky: fail fast after the first timeout
      if (clock64() - t0 > 120000000000LL) { atomicExch(err, 1); break; }      // ~60 s at 1.9 GHz
    }
  }
  return (v & 0xffffffULL) * (unsigned long long)kPeerGranule;
}
__device__ __forceinline__ void peerSignal(const PeerMsg& m, unsigned long long k) {
  peerStoreFlag(m.remoteFlag, (k << 24) | m.remoteOff[k & 1]);
}
#endif

}  // namespace fvmgpu
