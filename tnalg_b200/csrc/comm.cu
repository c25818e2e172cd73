// In-library communicator (SURVEY.md 8b `tn_comm_init(ncclComm_t)`, 8e): the collectives of the sharded hot path are issued
// by the library itself, stream-ordered, instead of through a Python callback per matvec.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a process that already carries a libnccl (e.g. the one bundled with
// torch) shares it, so a communicator created by the caller can be adopted (tn_comm_init) and one created here
// (tn_comm_init_rank, unique id exchanged by the caller's own bootstrap) lives next to torch.distributed's.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <mutex>
#include <vector>

#include "comm.cuh"

namespace tn {

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // RTLD_NOLOAD first: reuse the libnccl the process already has (same soname), else load the system one
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.handle = h;
#define TN_SYM(name) *reinterpret_cast<void**>(&api.name) = dlsym(h, "nccl" #name)
    TN_SYM(GetUniqueId); TN_SYM(CommInitRank); TN_SYM(CommDestroy); TN_SYM(AllReduce); TN_SYM(AllGather); TN_SYM(Broadcast);
    TN_SYM(GroupStart); TN_SYM(GroupEnd); TN_SYM(GetErrorString);
#undef TN_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.Broadcast &&
             api.GroupStart && api.GroupEnd && api.GetErrorString;
  });
  return api.ok ? &api : nullptr;
}
}  // namespace

#define TN_NCCL(api, x)                                                                      \
  do {                                                                                       \
    ncclResult_t r_ = (x);                                                                   \
    if (r_ != ncclSuccess) {                                                                 \
      tn::set_error("%s failed: %s (%s:%d)", #x, (api)->GetErrorString(r_), __FILE__, __LINE__); \
      return TN_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------------
// Peer window (NVLink / NVSwitch peer memory, CUDA IPC between the processes of one box)
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr size_t kArFlagOff = 0;       // u64 ar_flag[kPeerRing][kPeerMaxWorld]: sequence number of the contribution in the slot
constexpr size_t kAgCountOff = 1024;   // u64 ag_count[kPeerMaxWorld]: CTA arrivals signalled by every source so far
constexpr size_t kArSlotOff = 4096;    // double ar_slot[kPeerRing][kPeerMaxWorld][kPeerArMax]
static_assert(kArSlotOff + sizeof(double) * kPeerRing * kPeerMaxWorld * kPeerArMax <= kPeerHeaderBytes, "peer window header layout");
constexpr long long kSpinLimit = 40000000000LL;   // ~20 s of clock64() at 2 GHz: a lost peer aborts the kernel instead of hanging the GPU

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void spin_until_at_least(const unsigned long long* p, unsigned long long want) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < want) {
    if (clock64() - t0 > kSpinLimit) __trap();
  }
}

// buf[i] <- sum over ranks of buf[i], i < count <= kPeerArMax: every rank stores its contribution into slot [ring][rank] of
// EVERY window, raises the matching flag there, waits for the world's flags in its own window and adds the slots in rank
// order -- the same order on every rank, so the result is bit-identical everywhere.  One CTA, no NCCL launch.
// pred != nullptr && *pred == 0: the call is skipped -- on EVERY rank, because the flag must derive from all-reduced values
// (sequence numbers only grow, so a skipped slot is simply never waited for).
__global__ void peer_allreduce_kernel(PeerView pv, double* __restrict__ buf, int count, unsigned long long seq,
                                      const int* __restrict__ pred) {
  if (pred && *pred == 0) return;
  const int tid = threadIdx.x;
  const int ring = (int)(seq % kPeerRing);
  if (tid < count) {
    const double v = buf[tid];
    for (int dst = 0; dst < pv.world; ++dst)
      reinterpret_cast<double*>(pv.win[dst] + kArSlotOff)[((size_t)ring * kPeerMaxWorld + pv.rank) * kPeerArMax + tid] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < pv.world)
    st_release_sys(reinterpret_cast<unsigned long long*>(pv.win[tid] + kArFlagOff) + ring * kPeerMaxWorld + pv.rank, seq);
  if (tid < pv.world)
    spin_until_at_least(reinterpret_cast<const unsigned long long*>(pv.win[pv.rank] + kArFlagOff) + ring * kPeerMaxWorld + tid, seq);
  __syncthreads();
  if (tid < count) {
    const double* slot = reinterpret_cast<const double*>(pv.win[pv.rank] + kArSlotOff) + (size_t)ring * kPeerMaxWorld * kPeerArMax;
    double s = 0.0;
    for (int src = 0; src < pv.world; ++src) s += ld_relaxed_sys(slot + (size_t)src * kPeerArMax + tid);
    buf[tid] = s;
  }
}

// all-gather, push half: this rank's slice goes to byte offset `off` of every window (its own included) with 16-byte stores;
// every CTA then adds one arrival to counter [rank] of every window
__global__ void __launch_bounds__(256) peer_push_kernel(PeerView pv, const double* __restrict__ src, long long count, size_t off) {
  const long long pairs = count / 2;
  const double2* s2 = reinterpret_cast<const double2*>(src);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (long long)gridDim.x * blockDim.x) {
    const double2 v = s2[i];
    for (int dst = 0; dst < pv.world; ++dst) reinterpret_cast<double2*>(pv.win[dst] + off)[i] = v;
  }
  if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const double v = src[count - 1];
    for (int dst = 0; dst < pv.world; ++dst) reinterpret_cast<double*>(pv.win[dst] + off)[count - 1] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < pv.world)
    red_release_sys_add(reinterpret_cast<unsigned long long*>(pv.win[threadIdx.x] + kAgCountOff) + pv.rank, 1ULL);
}
// 8-byte variant for slices that are not 16-byte aligned
__global__ void __launch_bounds__(256) peer_push_kernel_f64(PeerView pv, const double* __restrict__ src, long long count, size_t off) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const double v = src[i];
    for (int dst = 0; dst < pv.world; ++dst) reinterpret_cast<double*>(pv.win[dst] + off)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < pv.world)
    red_release_sys_add(reinterpret_cast<unsigned long long*>(pv.win[threadIdx.x] + kAgCountOff) + pv.rank, 1ULL);
}
// all-gather, wait half: returns once every source has signalled `expected` arrivals in this rank's window
__global__ void peer_wait_kernel(PeerView pv, unsigned long long expected) {
  if (threadIdx.x < pv.world)
    spin_until_at_least(reinterpret_cast<const unsigned long long*>(pv.win[pv.rank] + kAgCountOff) + threadIdx.x, expected);
}

int nccl_allreduce(tn_comm* c, double* buf, long long count, cudaStream_t stream) {
  NcclApi* api = nccl_api();
  TN_REQUIRE(api && c && c->nccl, "tn_comm: no communicator");
  TN_NCCL(api, api->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, static_cast<ncclComm_t>(c->nccl), stream));
  return TN_OK;
}

void peer_window_release(tn_comm* c, bool free_local) {
  PeerWindow* pw = c->pw;
  if (!pw || pw->state != 1) return;
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && pw->view.win[r]) cudaIpcCloseMemHandle(pw->view.win[r]);
  if (free_local && pw->view.win[c->rank]) cudaFree(pw->view.win[c->rank]);
  for (int r = 0; r < kPeerMaxWorld; ++r) pw->view.win[r] = nullptr;
  pw->state = 0;
  pw->bytes = 0;
}

// collective: every rank calls it with the same `need` at the same point of its stream
PeerWindow* peer_window_ensure(tn_comm* c, size_t need, cudaStream_t stream) {
  if (!c || c->world < 2 || c->world > kPeerMaxWorld) return nullptr;
  if (!c->pw) {
    c->pw = new PeerWindow();
    if (getenv("TNALG_NO_PEER")) c->pw->state = -1;
  }
  PeerWindow* pw = c->pw;
  if (pw->state < 0) return nullptr;
  if (pw->state == 1 && need <= pw->bytes) return pw;
  NcclApi* api = nccl_api();
  if (!api || !c->nccl) return nullptr;
  const size_t gran = (size_t)2 << 20;
  const size_t newbytes = (std::max(need, pw->state == 1 ? 2 * pw->bytes : (size_t)0) + gran - 1) / gran * gran;
  auto barrier = [&]() -> bool {   // NCCL all-reduce of one double + stream sync
    return api->AllReduce(pw->staging, pw->staging, 1, ncclDouble, ncclSum, static_cast<ncclComm_t>(c->nccl), stream) == ncclSuccess &&
           cudaStreamSynchronize(stream) == cudaSuccess;
  };
  if (cudaStreamSynchronize(stream) != cudaSuccess) return nullptr;
  if (!pw->staging && cudaMalloc(&pw->staging, 64 * (kPeerMaxWorld + 1)) != cudaSuccess) {
    cudaGetLastError();
    pw->state = -1;
    return nullptr;
  }
  if (pw->state == 1) {   // grow: nobody may still address the old windows when they are freed
    const int rank = c->rank;
    for (int r = 0; r < c->world; ++r)
      if (r != rank && pw->view.win[r]) cudaIpcCloseMemHandle(pw->view.win[r]);
    char* old_local = pw->view.win[rank];
    cudaMemset(pw->staging, 0, sizeof(double));
    barrier();
    cudaFree(old_local);
    for (int r = 0; r < kPeerMaxWorld; ++r) pw->view.win[r] = nullptr;
    pw->state = 0;
  }
  char* local = nullptr;
  bool ok = cudaMalloc(&local, newbytes) == cudaSuccess;
  if (ok) ok = cudaMemset(local, 0, kPeerHeaderBytes) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  if (ok) ok = cudaIpcGetMemHandle(&mine, local) == cudaSuccess;
  char* stage = static_cast<char*>(pw->staging);
  std::vector<cudaIpcMemHandle_t> all((size_t)c->world);
  bool xok = cudaMemcpy(stage + 64 * c->rank, &mine, 64, cudaMemcpyHostToDevice) == cudaSuccess &&
             api->AllGather(stage + 64 * c->rank, stage, 64, ncclChar, static_cast<ncclComm_t>(c->nccl), stream) == ncclSuccess &&
             cudaStreamSynchronize(stream) == cudaSuccess &&
             cudaMemcpy(all.data(), stage, 64 * (size_t)c->world, cudaMemcpyDeviceToHost) == cudaSuccess;
  PeerView view{};
  view.rank = c->rank;
  view.world = c->world;
  if (ok && xok) {
    view.win[c->rank] = local;
    for (int r = 0; r < c->world && ok; ++r) {
      if (r == c->rank) continue;
      void* ptr = nullptr;
      ok = cudaIpcOpenMemHandle(&ptr, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      view.win[r] = static_cast<char*>(ptr);
    }
  }
  cudaGetLastError();   // a failed IPC call must not poison later launches
  // agreement: the window is used only if EVERY rank mapped every peer
  double fails = (ok && xok) ? 0.0 : 1.0;
  bool agreed = cudaMemcpy(stage + 64 * kPeerMaxWorld, &fails, sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
                api->AllReduce(stage + 64 * kPeerMaxWorld, stage + 64 * kPeerMaxWorld, 1, ncclDouble, ncclSum,
                               static_cast<ncclComm_t>(c->nccl), stream) == ncclSuccess &&
                cudaStreamSynchronize(stream) == cudaSuccess &&
                cudaMemcpy(&fails, stage + 64 * kPeerMaxWorld, sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
  if (!agreed || fails != 0.0) {
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && view.win[r]) cudaIpcCloseMemHandle(view.win[r]);
    if (local) cudaFree(local);
    cudaGetLastError();
    pw->state = -1;
    if (c->rank == 0) fprintf(stderr, "tnalg_b200: peer window unavailable (CUDA IPC mapping failed on %d rank(s)); NCCL collectives are used\n", (int)fails);
    return nullptr;
  }
  pw->view = view;
  pw->bytes = newbytes;
  pw->ar_seq = pw->ag_seq = pw->ag_arrivals = 0;
  pw->state = 1;
  return pw;
}
}  // namespace

bool comm_peer_available(tn_comm* c, cudaStream_t stream) { return c && peer_window_ensure(c, kPeerHeaderBytes, stream) != nullptr; }

int comm_allreduce_sum(tn_comm* c, double* buf, long long count, cudaStream_t stream, const int* pred) {
  TN_REQUIRE(c, "tn_comm: no communicator");
  if (count <= kPeerArMax) {
    PeerWindow* pw = peer_window_ensure(c, kPeerHeaderBytes, stream);
    if (pw) {
      ++pw->ar_seq;
      peer_allreduce_kernel<<<1, kPeerArMax, 0, stream>>>(pw->view, buf, (int)count, pw->ar_seq, pred);
      TN_LAUNCHED();
      ++c->n_collectives;
      ++c->n_peer_collectives;
      return TN_OK;
    }
  }
  TN_REQUIRE(!pred, "tn_comm: a predicated all-reduce needs the peer window");
  TN_CHECK(nccl_allreduce(c, buf, count, stream));
  ++c->n_collectives;
  return TN_OK;
}

int comm_allgather(tn_comm* c, const double* send, double* recv, long long count_per_rank, cudaStream_t stream) {
  NcclApi* api = nccl_api();
  TN_REQUIRE(api && c && c->nccl, "tn_comm: no communicator");
  TN_NCCL(api, api->AllGather(send, recv, (size_t)count_per_rank, ncclDouble, static_cast<ncclComm_t>(c->nccl), stream));
  ++c->n_collectives;
  return TN_OK;
}

// Two data regions used alternately: a rank may push the next vector while a slower peer still reads the previous one.  A
// region is rewritten only two all-gathers later; the callers (sharded Lanczos) have an all-reduce -- a barrier -- between
// consecutive all-gathers, so every peer has finished reading it by then.
int comm_allgather_window(tn_comm* c, const double* send, long long count_per_rank, double* recv_fallback, const double** out,
                          cudaStream_t stream) {
  TN_REQUIRE(c && send && out && count_per_rank > 0, "tn_comm: bad all-gather arguments");
  const size_t region = align_up(sizeof(double) * (size_t)count_per_rank * (size_t)c->world);
  PeerWindow* pw = peer_window_ensure(c, kPeerHeaderBytes + 2 * region, stream);
  if (!pw) {
    TN_REQUIRE(recv_fallback, "tn_comm: all-gather without a receive buffer");
    TN_CHECK(comm_allgather(c, send, recv_fallback, count_per_rank, stream));
    *out = recv_fallback;
    return TN_OK;
  }
  const size_t half = ((pw->bytes - kPeerHeaderBytes) / 2) / 256 * 256;
  const size_t base = kPeerHeaderBytes + (size_t)(pw->ag_seq & 1) * half;
  ++pw->ag_seq;
  const size_t off = base + sizeof(double) * (size_t)count_per_rank * (size_t)c->rank;
  const bool vec2 = (reinterpret_cast<uintptr_t>(send) & 15) == 0 && (off & 15) == 0;
  // the grid is a function of the slice length alone: every rank counts `grid` arrivals per source, while the 16-byte path
  // depends on this rank's offset (odd slice lengths put odd ranks on the 8-byte kernel)
  const int grid = (int)std::max<long long>(1, std::min<long long>((count_per_rank + 511) / 512, 2LL * sm_count()));
  if (vec2) peer_push_kernel<<<grid, 256, 0, stream>>>(pw->view, send, count_per_rank, off);
  else peer_push_kernel_f64<<<grid, 256, 0, stream>>>(pw->view, send, count_per_rank, off);
  TN_LAUNCHED();
  pw->ag_arrivals += (unsigned long long)grid;
  peer_wait_kernel<<<1, 32, 0, stream>>>(pw->view, pw->ag_arrivals);
  TN_LAUNCHED();
  ++c->n_collectives;
  ++c->n_peer_collectives;
  *out = reinterpret_cast<const double*>(pw->view.win[c->rank] + base);
  return TN_OK;
}

int comm_broadcast_many(tn_comm* c, double* const* bufs, const long long* counts, const int* roots, int n, cudaStream_t stream) {
  NcclApi* api = nccl_api();
  TN_REQUIRE(api && c && c->nccl, "tn_comm: no communicator");
  if (n <= 0) return TN_OK;
  TN_NCCL(api, api->GroupStart());
  for (int i = 0; i < n; ++i) {
    ncclResult_t r = api->Broadcast(bufs[i], bufs[i], (size_t)counts[i], ncclDouble, roots[i], static_cast<ncclComm_t>(c->nccl), stream);
    if (r != ncclSuccess) {
      api->GroupEnd();
      set_error("ncclBroadcast failed: %s", api->GetErrorString(r));
      return TN_ERR_CUDA;
    }
  }
  TN_NCCL(api, api->GroupEnd());
  ++c->n_collectives;
  return TN_OK;
}

}  // namespace tn

using namespace tn;

extern "C" int tn_comm_unique_id(char* id128) {
  TN_REQUIRE(id128, "tn_comm_unique_id: null buffer");
  NcclApi* api = nccl_api();
  TN_REQUIRE(api, "tn_comm: libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbols");
  ncclUniqueId id;
  TN_NCCL(api, api->GetUniqueId(&id));
  static_assert(sizeof(id) == 128, "ncclUniqueId size");
  memcpy(id128, &id, 128);
  return TN_OK;
}

extern "C" int tn_comm_init_rank(tn_comm** out, const char* id128, int rank, int world) {
  TN_REQUIRE(out && id128 && world >= 1 && rank >= 0 && rank < world, "tn_comm_init_rank: bad arguments");
  NcclApi* api = nccl_api();
  TN_REQUIRE(api, "tn_comm: libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  TN_NCCL(api, api->CommInitRank(&comm, world, id, rank));
  tn_comm* c = new tn_comm();
  c->nccl = comm; c->rank = rank; c->world = world; c->owned = 1; c->n_collectives = 0; c->n_peer_collectives = 0; c->pw = nullptr;
  *out = c;
  return TN_OK;
}

extern "C" int tn_comm_init(tn_comm** out, void* nccl_comm, int rank, int world) {
  TN_REQUIRE(out && nccl_comm && world >= 1 && rank >= 0 && rank < world, "tn_comm_init: bad arguments");
  TN_REQUIRE(nccl_api(), "tn_comm: libnccl.so.2 could not be loaded");
  tn_comm* c = new tn_comm();
  c->nccl = nccl_comm; c->rank = rank; c->world = world; c->owned = 0; c->n_collectives = 0; c->n_peer_collectives = 0; c->pw = nullptr;
  *out = c;
  return TN_OK;
}

extern "C" int tn_comm_rank(const tn_comm* c) { return c ? c->rank : -1; }
extern "C" int tn_comm_world(const tn_comm* c) { return c ? c->world : 0; }
extern "C" long long tn_comm_collectives(const tn_comm* c) { return c ? c->n_collectives : 0; }
extern "C" long long tn_comm_peer_collectives(const tn_comm* c) { return c ? c->n_peer_collectives : 0; }

extern "C" int tn_comm_destroy(tn_comm* c) {
  if (!c) return TN_OK;
  NcclApi* api = nccl_api();
  if (c->pw) {
    cudaDeviceSynchronize();
    peer_window_release(c, true);
    if (c->pw->staging) cudaFree(c->pw->staging);
    delete c->pw;
  }
  if (c->owned && api && c->nccl) api->CommDestroy(static_cast<ncclComm_t>(c->nccl));
  delete c;
  return TN_OK;
}

extern "C" int tn_comm_allreduce_sum(tn_comm* c, double* buf, long long count, void* stream) {
  TN_REQUIRE(buf && count > 0, "tn_comm_allreduce_sum: bad arguments");
  return comm_allreduce_sum(c, buf, count, static_cast<cudaStream_t>(stream));
}

extern "C" int tn_comm_allgather(tn_comm* c, const double* send, double* recv, long long count_per_rank, void* stream) {
  TN_REQUIRE(send && recv && count_per_rank > 0, "tn_comm_allgather: bad arguments");
  return comm_allgather(c, send, recv, count_per_rank, static_cast<cudaStream_t>(stream));
}

extern "C" int tn_comm_broadcast_many(tn_comm* c, double* const* bufs, const long long* counts, const int* roots, int n,
                                      void* stream) {
  TN_REQUIRE(n == 0 || (bufs && counts && roots), "tn_comm_broadcast_many: null arrays");
  for (int i = 0; i < n; ++i) TN_REQUIRE(bufs[i] && counts[i] > 0 && roots[i] >= 0 && c && roots[i] < c->world, "tn_comm_broadcast_many: bad entry %d", i);
  return comm_broadcast_many(c, bufs, counts, roots, n, static_cast<cudaStream_t>(stream));
}
