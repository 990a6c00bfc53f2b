/*
 * nw_comm.h -- NCCL communicator, bound at run time with dlopen so that the
 * library loads (and its host logic can be tested) on machines without NCCL
 * or without a GPU.  Replaces the MPI communicator used by STK / hypre on this
 * path (stk::mesh::parallel_sum, HYPRE_IJMatrixAssemble).
 */
#ifndef NW_COMM_H
#define NW_COMM_H

#include <cuda_runtime.h>

#include <string>

#include "nw_internal.h"

namespace nw {

bool comm_unique_id(void* out128, std::string& err);
bool comm_init(
  Comm& c, const void* uniqueId, int nranks, int rank, std::string& err);
void comm_destroy(Comm& c);

/* grouped neighbour exchange: for every peer i send sendCount[i] doubles from
 * sendPtr[i] and receive recvCount[i] doubles into recvPtr[i] */
bool comm_exchange_f64(
  Comm& c, int nPeers, const int* peers, const double* const* sendPtr,
  const int64_t* sendCount, double* const* recvPtr, const int64_t* recvCount,
  cudaStream_t s, std::string& err);
bool comm_exchange_i64(
  Comm& c, int nPeers, const int* peers, const int64_t* const* sendPtr,
  const int64_t* sendCount, int64_t* const* recvPtr, const int64_t* recvCount,
  cudaStream_t s, std::string& err);
/* in-place sum of n doubles over all ranks */
bool comm_allreduce_sum_f64(
  Comm& c, double* buf, int64_t n, cudaStream_t s, std::string& err);

} // namespace nw

#endif
