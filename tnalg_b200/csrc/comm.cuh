// Internal view of the in-library communicator (comm.cu).
#pragma once
#include <cstring>

#include "common.cuh"

namespace tn {
constexpr int kPeerMaxWorld = 8;          // GPUs of one NVSwitch box
constexpr int kPeerRing = 4;              // all-reduce slots in flight: an executed all-reduce is a barrier, and callers never skip
                                          // (predicate) more than one in a row, so a slot is rewritten only after a later barrier
constexpr int kPeerArMax = 64;            // doubles per small all-reduce
constexpr size_t kPeerHeaderBytes = 65536;  // flags + all-reduce slots in front of the data region

// Peer window: one cudaMalloc'ed buffer per rank, mapped into every other process of the box with CUDA IPC, so that kernels
// exchange data with plain stores over NVLink instead of NCCL launches (the sharded Lanczos step is latency bound: two
// all-reduces of ~22 numbers and one all-gather of a 2 MB slice per step).  Device-side view, passed to kernels by value.
struct PeerView {
  char* win[kPeerMaxWorld];   // win[r]: rank r's window in this process' address space (win[rank] is the local one)
  int rank, world;
};
struct PeerWindow {
  int state = 0;              // 0: not tried, 1: enabled, -1: unavailable (IPC mapping failed on some rank) -> NCCL path
  size_t bytes = 0;           // size of every rank's window
  PeerView view{};
  unsigned long long ar_seq = 0;        // small all-reduces issued (same on every rank)
  unsigned long long ag_seq = 0;        // all-gathers issued
  unsigned long long ag_arrivals = 0;   // CTA arrivals every source has signalled so far (same on every rank)
  void* staging = nullptr;    // device buffer for the handle exchange
};
}  // namespace tn

struct tn_comm {
  void* nccl;  // ncclComm_t
  int rank, world;
  int owned;   // created by tn_comm_init_rank (destroyed with the handle) vs adopted from the caller
  long long n_collectives;
  long long n_peer_collectives;   // ... of which went through the peer window
  tn::PeerWindow* pw;
};

namespace tn {
// pred (device flag, identical on every rank because it derives from all-reduced values): 0 skips the call everywhere; only
// with the peer window (comm_peer_available) and count <= kPeerArMax
int comm_allreduce_sum(tn_comm* c, double* buf, long long count, cudaStream_t stream, const int* pred = nullptr);
// collective (may set the window up): true when small all-reduces run as peer-window kernels
bool comm_peer_available(tn_comm* c, cudaStream_t stream);
// equal contributions: recv holds world * count_per_rank doubles, rank r's block at r * count_per_rank
int comm_allgather(tn_comm* c, const double* send, double* recv, long long count_per_rank, cudaStream_t stream);
// same result as comm_allgather, but the gathered vector lives in the library's peer window when that is available (*out points
// to it; valid until the all-gather after next); falls back to NCCL into `recv_fallback`
int comm_allgather_window(tn_comm* c, const double* send, long long count_per_rank, double* recv_fallback, const double** out,
                          cudaStream_t stream);
// one grouped launch of n broadcasts (in place)
int comm_broadcast_many(tn_comm* c, double* const* bufs, const long long* counts, const int* roots, int n, cudaStream_t stream);
}  // namespace tn
