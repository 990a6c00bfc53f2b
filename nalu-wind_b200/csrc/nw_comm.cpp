/*
 * nw_comm.cpp -- NCCL through dlopen (see nw_comm.h).
 */
#include "nw_comm.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace nw {

namespace {

struct UniqueId
{
  char internal[128];
};
typedef int ncclResult_t;
typedef void* ncclComm_t;
enum { kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0 };

struct Api
{
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(UniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, UniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) =
    nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) =
    nullptr;
  ncclResult_t (*AllReduce)(
    const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string loadError;
};

Api&
api()
{
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    /* prefer a libnccl that is already in the process (e.g. torch's) */
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if (a.handle)
        break;
    }
    if (!a.handle)
      for (const char* n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle)
          break;
      }
    if (!a.handle) {
      a.loadError = std::string("cannot load libnccl: ") + dlerror();
      return;
    }
#define NW_SYM(field, name)                                           \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name)); \
  if (!a.field)                                                       \
    a.loadError = std::string("libnccl lacks ") + name;
    NW_SYM(GetUniqueId, "ncclGetUniqueId")
    NW_SYM(CommInitRank, "ncclCommInitRank")
    NW_SYM(CommDestroy, "ncclCommDestroy")
    NW_SYM(GroupStart, "ncclGroupStart")
    NW_SYM(GroupEnd, "ncclGroupEnd")
    NW_SYM(Send, "ncclSend")
    NW_SYM(Recv, "ncclRecv")
    NW_SYM(AllReduce, "ncclAllReduce")
    NW_SYM(GetErrorString, "ncclGetErrorString")
#undef NW_SYM
  });
  return a;
}

bool
check(ncclResult_t r, const char* what, std::string& err)
{
  if (r == 0)
    return true;
  err = std::string(what) + ": " + api().GetErrorString(r);
  return false;
}

template <class T>
bool
exchange(
  Comm& c,
  int dtype,
  int nPeers,
  const int* peers,
  const T* const* sendPtr,
  const int64_t* sendCount,
  T* const* recvPtr,
  const int64_t* recvCount,
  cudaStream_t s,
  std::string& err)
{
  Api& a = api();
  if (!a.loadError.empty() || !c.comm) {
    err = a.loadError.empty() ? "communicator not initialised" : a.loadError;
    return false;
  }
  if (!check(a.GroupStart(), "ncclGroupStart", err))
    return false;
  /* a failed send / receive must not leave the group open (the communicator
   * would be unusable afterwards): always close it, report the first error */
  bool ok = true;
  for (int i = 0; i < nPeers && ok; ++i) {
    if (sendCount[i] > 0)
      ok = check(
        a.Send(sendPtr[i], (size_t)sendCount[i], dtype, peers[i], c.comm, s),
        "ncclSend", err);
    if (ok && recvCount[i] > 0)
      ok = check(
        a.Recv(recvPtr[i], (size_t)recvCount[i], dtype, peers[i], c.comm, s),
        "ncclRecv", err);
  }
  std::string endErr;
  const bool ended = check(a.GroupEnd(), "ncclGroupEnd", endErr);
  if (ok && !ended)
    err = endErr;
  return ok && ended;
}

} // namespace

bool
comm_unique_id(void* out128, std::string& err)
{
  Api& a = api();
  if (!a.loadError.empty()) {
    err = a.loadError;
    return false;
  }
  UniqueId id;
  if (!check(a.GetUniqueId(&id), "ncclGetUniqueId", err))
    return false;
  std::memcpy(out128, &id, sizeof(id));
  return true;
}

bool
comm_init(Comm& c, const void* uniqueId, int nranks, int rank, std::string& err)
{
  Api& a = api();
  if (!a.loadError.empty()) {
    err = a.loadError;
    return false;
  }
  UniqueId id;
  std::memcpy(&id, uniqueId, sizeof(id));
  ncclComm_t comm = nullptr;
  if (!check(a.CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank", err))
    return false;
  c.comm = comm;
  c.nranks = nranks;
  c.rank = rank;
  return true;
}

void
comm_destroy(Comm& c)
{
  if (c.comm && api().CommDestroy)
    api().CommDestroy(c.comm);
  c.comm = nullptr;
}

bool
comm_exchange_f64(
  Comm& c,
  int nPeers,
  const int* peers,
  const double* const* sendPtr,
  const int64_t* sendCount,
  double* const* recvPtr,
  const int64_t* recvCount,
  cudaStream_t s,
  std::string& err)
{
  return exchange<double>(
    c, kNcclFloat64, nPeers, peers, sendPtr, sendCount, recvPtr, recvCount, s,
    err);
}

bool
comm_exchange_i64(
  Comm& c,
  int nPeers,
  const int* peers,
  const int64_t* const* sendPtr,
  const int64_t* sendCount,
  int64_t* const* recvPtr,
  const int64_t* recvCount,
  cudaStream_t s,
  std::string& err)
{
  return exchange<int64_t>(
    c, kNcclInt64, nPeers, peers, sendPtr, sendCount, recvPtr, recvCount, s,
    err);
}

bool
comm_allreduce_sum_f64(
  Comm& c, double* buf, int64_t n, cudaStream_t s, std::string& err)
{
  Api& a = api();
  if (!a.loadError.empty() || !c.comm) {
    err = a.loadError.empty() ? "communicator not initialised" : a.loadError;
    return false;
  }
  return check(
    a.AllReduce(buf, buf, (size_t)n, kNcclFloat64, kNcclSum, c.comm, s),
    "ncclAllReduce", err);
}

} // namespace nw
