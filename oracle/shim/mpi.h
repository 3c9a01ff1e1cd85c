// oracle shim (test infrastructure only): single-rank stand-in for the MPI-2 C++ bindings
// the reference uses under -DFVM_PARALLEL (F/MultiField.cpp:488-551, F/AMG.cpp:165,
// F/MultiFieldReduction.cpp:213-225 ...). No MPI runtime exists in this image.
// -DFVM_PARALLEL is needed because the reference's goldens were produced by the parallel
// build (coarse level push/break order, F/AMG.cpp:171-180 vs :199-204).
// With one rank every collective is an identity / memcpy and no point-to-point is reached.
#pragma once
#include <cstring>
#include <cstdlib>
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
namespace MPI {
struct Datatype { int size; };
static const Datatype INT = {4}, BYTE = {1}, DOUBLE = {8}, CHAR = {1}, BOOL = {1}, DOUBLE_INT = {16};
struct Op { int id; };
static const Op SUM = {0}, MIN = {1}, MAX = {2}, PROD = {3}, MINLOC = {4}, MAXLOC = {5};
static void* const IN_PLACE = (void*)-1;
static const int MAX_PORT_NAME = 256;
struct Info {};
static const Info INFO_NULL = Info();
struct Status {
  int Get_source() const { return 0; }
  int Get_tag() const { return 0; }
  int Get_count(const Datatype&) const { return 0; }
};
struct Request {
  static void Waitall(int, Request*) {}
  void Wait() {}
};
struct Intercomm;
struct Comm {
  static Intercomm Get_parent();
};
struct Intracomm {
  int Get_rank() const { return 0; }
  int Get_size() const { return 1; }
  void Allreduce(const void* s, void* r, int n, const Datatype& t, const Op&) const {
    if (s != IN_PLACE) std::memcpy(r, s, (size_t)n * t.size);
  }
  void Reduce(const void* s, void* r, int n, const Datatype& t, const Op&, int) const {
    if (s != IN_PLACE) std::memcpy(r, s, (size_t)n * t.size);
  }
  void Bcast(void*, int, const Datatype&, int) const {}
  void Barrier() const {}
  void Abort(int) const { std::abort(); }
  Request Isend(const void*, int, const Datatype&, int, int) const { return Request(); }
  Request Irecv(void*, int, const Datatype&, int, int) const { return Request(); }
  void Send(const void*, int, const Datatype&, int, int) const {}
  void Recv(void*, int, const Datatype&, int, int) const {}
  void Recv(void*, int, const Datatype&, int, int, Status&) const {}
  void Sendrecv(const void* s, int ns, const Datatype& ts, int, int, void* r, int, const Datatype&, int,
                int) const {
    std::memcpy(r, s, (size_t)ns * ts.size);
  }
  void Sendrecv(const void* s, int ns, const Datatype& ts, int, int, void* r, int, const Datatype&, int, int,
                Status&) const {
    std::memcpy(r, s, (size_t)ns * ts.size);
  }
  bool Iprobe(int, int, Status&) const { return true; }
  bool Iprobe(int, int) const { return true; }
  Intracomm Split(int, int) const { return *this; }
  void Gather(const void* s, int ns, const Datatype& ts, void* r, int, const Datatype&, int) const {
    std::memcpy(r, s, (size_t)ns * ts.size);
  }
  void Allgather(const void* s, int ns, const Datatype& ts, void* r, int, const Datatype&) const {
    if (s != IN_PLACE) std::memcpy(r, s, (size_t)ns * ts.size);
  }
  void Gatherv(const void* s, int ns, const Datatype& ts, void* r, const int*, const int* displs,
               const Datatype&, int) const {
    std::memcpy((char*)r + (displs ? (size_t)displs[0] * ts.size : 0), s, (size_t)ns * ts.size);
  }
  void Allgatherv(const void* s, int ns, const Datatype& ts, void* r, const int*, const int* displs,
                  const Datatype&) const {
    if (s != IN_PLACE) std::memcpy((char*)r + (displs ? (size_t)displs[0] * ts.size : 0), s, (size_t)ns * ts.size);
  }
  void Scatterv(const void* s, const int*, const int* displs, const Datatype& ts, void* r, int nr,
                const Datatype& tr, int) const {
    std::memcpy(r, (const char*)s + (displs ? (size_t)displs[0] * ts.size : 0), (size_t)nr * tr.size);
  }
  Intercomm Connect(const char*, const Info&, int) const;
  void Free() {}
  operator MPI_Comm() const { return 0; }
  bool operator==(const Intracomm&) const { return true; }
  bool operator!=(const Intracomm&) const { return false; }
};
struct Intercomm : public Intracomm {
  Intracomm Merge(bool) const { return Intracomm(); }
};
inline Intercomm Comm::Get_parent() { return Intercomm(); }
inline Intercomm Intracomm::Connect(const char*, const Info&, int) const { return Intercomm(); }
static Intracomm COMM_WORLD;
static const Intracomm COMM_NULL = Intracomm();
inline void Init() {}
inline void Init(int&, char**&) {}
inline void Finalize() {}
inline bool Is_initialized() { return true; }
inline double Wtime() { return 0.0; }
struct Win {
  static Win Create(const void*, long, int, const Info&, const Intracomm&) { return Win(); }
  void Fence(int) const {}
  void Get(void*, int, const Datatype&, int, long, int, const Datatype&) const {}
  void Free() {}
};
typedef long Aint;
}  // namespace MPI
