// In-library communicator (SURVEY.md 8b `tn_comm_init(ncclComm_t)`, 8e): the collectives of the sharded hot path are issued
// by the library itself, stream-ordered, instead of through a Python callback per matvec.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a process that already carries a libnccl (e.g. the one bundled with
// torch) shares it, so a communicator created by the caller can be adopted (tn_comm_init) and one created here
// (tn_comm_init_rank, unique id exchanged by the caller's own bootstrap) lives next to torch.distributed's.
#include <dlfcn.h>
#include <nccl.h>

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

int comm_allreduce_sum(tn_comm* c, double* buf, long long count, cudaStream_t stream) {
  NcclApi* api = nccl_api();
  TN_REQUIRE(api && c && c->nccl, "tn_comm: no communicator");
  TN_NCCL(api, api->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, static_cast<ncclComm_t>(c->nccl), stream));
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
  c->nccl = comm; c->rank = rank; c->world = world; c->owned = 1; c->n_collectives = 0;
  *out = c;
  return TN_OK;
}

extern "C" int tn_comm_init(tn_comm** out, void* nccl_comm, int rank, int world) {
  TN_REQUIRE(out && nccl_comm && world >= 1 && rank >= 0 && rank < world, "tn_comm_init: bad arguments");
  TN_REQUIRE(nccl_api(), "tn_comm: libnccl.so.2 could not be loaded");
  tn_comm* c = new tn_comm();
  c->nccl = nccl_comm; c->rank = rank; c->world = world; c->owned = 0; c->n_collectives = 0;
  *out = c;
  return TN_OK;
}

extern "C" int tn_comm_rank(const tn_comm* c) { return c ? c->rank : -1; }
extern "C" int tn_comm_world(const tn_comm* c) { return c ? c->world : 0; }
extern "C" long long tn_comm_collectives(const tn_comm* c) { return c ? c->n_collectives : 0; }

extern "C" int tn_comm_destroy(tn_comm* c) {
  if (!c) return TN_OK;
  NcclApi* api = nccl_api();
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
