// Internal view of the in-library communicator (comm.cu).
#pragma once
#include <cstring>

#include "common.cuh"

struct tn_comm {
  void* nccl;  // ncclComm_t
  int rank, world;
  int owned;   // created by tn_comm_init_rank (destroyed with the handle) vs adopted from the caller
  long long n_collectives;
};

namespace tn {
int comm_allreduce_sum(tn_comm* c, double* buf, long long count, cudaStream_t stream);
// equal contributions: recv holds world * count_per_rank doubles, rank r's block at r * count_per_rank
int comm_allgather(tn_comm* c, const double* send, double* recv, long long count_per_rank, cudaStream_t stream);
// one grouped launch of n broadcasts (in place)
int comm_broadcast_many(tn_comm* c, double* const* bufs, const long long* counts, const int* roots, int n, cudaStream_t stream);
}  // namespace tn
