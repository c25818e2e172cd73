// Shared host/device helpers for libtnalg_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "tnalg_b200.h"

namespace tn {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
int sm_count();
// tn_set_deterministic(1): no stream-K split, no FP64-atomic combination of partial tiles, no concurrent stages -- every
// result is bit-reproducible from run to run (costs L2 operand sharing in the right matvec stage)
bool deterministic_mode();

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// carve aligned pieces out of a caller-provided workspace
struct Carver {
  char* base;
  size_t size, used;
  Carver(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
  template <class T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (used + bytes > size) return nullptr;
    T* r = reinterpret_cast<T*>(base + used);
    used += bytes;
    return r;
  }
};

#define TN_CUDA(x)                                                                                      \
  do {                                                                                                  \
    cudaError_t e_ = (x);                                                                               \
    if (e_ != cudaSuccess) {                                                                            \
      tn::set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__);          \
      return TN_ERR_CUDA;                                                                               \
    }                                                                                                   \
  } while (0)

#define TN_REQUIRE(cond, ...)         \
  do {                                \
    if (!(cond)) {                    \
      tn::set_error(__VA_ARGS__);     \
      return TN_ERR_INVALID;          \
    }                                 \
  } while (0)

#define TN_LAUNCHED()                                                                                   \
  do {                                                                                                  \
    ++tn::g_launches;                                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                                \
    if (e_ != cudaSuccess) {                                                                            \
      tn::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__);    \
      return TN_ERR_CUDA;                                                                               \
    }                                                                                                   \
  } while (0)

#define TN_CHECK(x)          \
  do {                       \
    int s_ = (x);            \
    if (s_ != TN_OK) return s_; \
  } while (0)

// k-block of every chain-GEMM kernel; the scheduler (ipl = ceil(K / kBK)), the cp.async kernel and the TMA kernel share it
#ifndef TN_BK
#define TN_BK 32
#endif
constexpr int kBK = TN_BK;
// CTA tile of the large configuration (both staging paths): 128x64 with two CTAs per SM by default
#ifndef TN_CFGL_BM
#define TN_CFGL_BM 128
#endif
#ifndef TN_CFGL_BN
#define TN_CFGL_BN 64
#endif
#ifndef TN_CFGL_CTAS
#define TN_CFGL_CTAS 2
#endif
#ifndef TN_STAGES_L
#define TN_STAGES_L 2
#endif
constexpr int kTileBM = TN_CFGL_BM, kTileBN = TN_CFGL_BN, kTileCtas = TN_CFGL_CTAS;

// ---- device-side descriptors of the chain GEMM (built by the host wrappers in chain_gemm.cu) ----
constexpr int kMaxD = TN_MAX_LOADPATH_DIM;  // operators applied inside the GEMM operand path
constexpr int kMaxPhys = TN_MAX_PHYS_DIM;   // element-wise site-operator kernels, effective-Hamiltonian plans
constexpr int kOpSlot = 24;   // doubles per operator slot in shared memory: kMaxD^2 entries + the has_op flag at [kOpFlag]
constexpr int kOpFlag = 16;
static_assert(kMaxD * kMaxD <= kOpFlag && kOpFlag < kOpSlot, "operator slot layout");

struct LinkDev {
  const double* A;
  const double* B;
  double op[kMaxD * kMaxD];
  int has_op;
  int a_dyn;  // 1: A = params.dyn_in (psi of this matvec call)
  int b_dyn;  // 1: B = params.dyn_in
  int a_map;  // TMA path: index of the operand's tensor map in the plan's map table (unused when a_dyn)
  int b_map;
  int pad;
};

struct ProblemDev {
  double* C;
  double alpha;
  int link_begin, link_count;
  int accumulate;
  int c_dyn;             // 1: C = params.dyn_out and alpha is multiplied by params.dyn_alpha
  int shared_out;        // 1: other problems of this launch add into the same C -> always combine with FP64 atomics
  int pad;
  long long work_begin;  // prefix sum of tiles*iters (stream-K) over the problems of the launch
  long long tile_begin;  // prefix sum of tiles
};

struct GemmParams {
  int M, N, K, d;
  int lda, ldb, ldc;
  int n_problems;
  int tiles_m, tiles_n;
  int ipl;  // k-iterations per link = ceil(K / BK)
  int split;  // 1: stream-K over [0,total_work) with atomics; 0: whole tiles, direct stores
  const ProblemDev* problems;
  const LinkDev* links;
  long long total_work, work_per_cta, total_tiles;
  const double* dyn_in;
  double* dyn_out;
  double dyn_alpha;
};

// internal launch entry shared by the plan / env-update / public wrappers
struct GemmLaunch {
  int mode, M, N, K, d, lda, ldb, ldc;
  int n_problems, n_links;
  int deterministic;
  int grid_div;  // > 1: this launch shares the SMs with grid_div - 1 concurrent launches and is scheduled on 1/grid_div of the CTA slots
};
// fills work_begin/tile_begin of `problems` (host copies), chooses tile config + split policy.
struct GemmSchedule {
  int config;  // 0 = 128x128 tiles, 1 = 64x64 tiles
  int aligned16;
  int split;
  int grid;
  int tiles_m, tiles_n, ipl;
  long long total_work, work_per_cta, total_tiles;
};
int gemm_plan_schedule(const GemmLaunch& L, ProblemDev* problems_host, const LinkDev* links_host, const double* dyn_in,
                       const double* dyn_out, GemmSchedule* S);
int gemm_launch(const GemmLaunch& L, const GemmSchedule& S, const ProblemDev* problems_dev, const LinkDev* links_dev,
                const double* dyn_in, double* dyn_out, double dyn_alpha, cudaStream_t stream);

}  // namespace tn
