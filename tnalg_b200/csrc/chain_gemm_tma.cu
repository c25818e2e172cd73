// TMA-staged variant of the chain GEMM (modes TN_NN and TN_NT) used by the effective-Hamiltonian matvec.
//
// Same math, tiling (128x64x32, 4 warps of 32x64, two CTAs per SM) and DMMA inner loop as chain_gemm.cu; the operand
// staging is done by the Tensor Memory Accelerator instead of per-thread cp.async:
//   * one elected thread issues 4-6 `cp.async.bulk.tensor` box loads (16 doubles = 128 B wide) per k-block; the boxes land
//     in shared memory in the SWIZZLE_128B_ATOM_32B layout: the 32-byte atom `a` of row `r` sits at position a ^ (r & 3)
//     (measured with tools/microbench/tma_probe.cu), which makes every LDS.64 fragment read conflict-free for both tile
//     orientations without padding;
//   * arrival of a stage is tracked by a `full` mbarrier (complete_tx), its release by a shared-memory arrival counter: the
//     last consumer warp to finish a stage refills it at once; there is no CTA-wide barrier inside the k loop;
//   * ragged edges (M, N, K not multiples of the tile) are zero-filled by the TMA unit itself;
//   * the (s, y) column grouping of the left stage is a 3-D tensor map over psi viewed as (a', s, b): the physical-index
//     "reshape" is just a coordinate of the box load.
// Tensor maps of the environment matrices are encoded once per plan (host, cuTensorMapEncodeTiled through the runtime's
// driver entry point) and live in the plan workspace; the two maps of psi are encoded per call and passed as
// __grid_constant__ kernel parameters.
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "tma.cuh"

namespace tn {

namespace {
#ifndef TN_TMA_STAGES
#define TN_TMA_STAGES TN_STAGES_L
#endif
constexpr int BM = kTileBM, BN = kTileBN, BK = kBK, WM = 32, WN = 64, STAGES = TN_TMA_STAGES;
constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN, THREADS = WARPS_M * WARPS_N * 32;
static_assert(BM == 128 && BN % WN == 0, "TMA kernel: unsupported tile");
constexpr int MT = WM / 8, NT = WN / 8;
constexpr int BOXW = 16;                      // doubles per box row (128 bytes)
constexpr int A_ELEMS = BM * BK;              // two boxes of BM x 16
constexpr int B_ELEMS = BN * BK;              // NT: two boxes of BN x 16; NN: four boxes of BK x 16
constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
constexpr unsigned STAGE_BYTES = STAGE_ELEMS * sizeof(double);
constexpr int kRing = STAGES + 2;  // operator ring: one slot per k-block in flight
constexpr size_t SMEM_BYTES = size_t(STAGE_ELEMS) * STAGES * sizeof(double) + kRing * kOpSlot * sizeof(double) + 2 * STAGES * sizeof(uint64_t);

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nTN_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra TN_WAIT_%=;\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
}  // namespace

// psi_map_a: psi as the (M = a*d) x (K = b) A operand of the right stage (2-D, box 16 x BM)
// psi_map_b: psi as the (K = a') x (s) x (y = b) B operand of the left stage (3-D, box 16 x 1 x BK)
template <int MODE, int DMAX>
__global__ void __launch_bounds__(THREADS, kTileCtas) chain_gemm_tma_kernel(const GemmParams p, const CUtensorMap* __restrict__ maps,
                                                                    const __grid_constant__ CUtensorMap psi_map_a,
                                                                    const __grid_constant__ CUtensorMap psi_map_b) {
  extern __shared__ __align__(1024) double smem[];  // the swizzle pattern is a function of the shared address: 1024-byte aligned base
  if (smem_u32(smem) & 1023u) __trap();
  double* const sOpRing = smem + STAGE_ELEMS * STAGES;
  uint64_t* const full = reinterpret_cast<uint64_t*>(sOpRing + kRing * kOpSlot);
  int* const done = reinterpret_cast<int*>(full + STAGES);  // per stage: consumer warps that finished reading it

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int warp_m = warp % WARPS_M, warp_n = warp / WARPS_M;
  const int d = p.d;
  const int BNy = (MODE == TN_NN && d > 1) ? ((BN / d) / BOXW) * BOXW : BN;  // y values per tile (multiple of the box width)
  const int BMe = (MODE == TN_NT && d > 1) ? (BM / d) * d : BM;
  const int Ny = (MODE == TN_NN && d > 1) ? p.N / d : p.N;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      done[2 * s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // fragment <-> tile mapping
  int m_s[MT], m_base[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int r = warp_m * WM + mt * 8 + g;
    m_s[mt] = (MODE == TN_NT && d > 1) ? r % d : 0;
    m_base[mt] = r - m_s[mt];
  }
  int n_s[NT], n_y[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int c = warp_n * WN + nt * 8 + g;
    const int s = (MODE == TN_NN && d > 1) ? c / BNy : 0;
    n_y[nt] = (MODE == TN_NN && d > 1) ? c % BNy : c;
    n_s[nt] = s < d ? s : 0;
  }
  int ep_s[NT], ep_y[NT];  // epilogue: (s, y offset) of fragment columns 2t, 2t+1 of each n-tile (NN with an operator grouping)
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int c = warp_n * WN + nt * 8 + 2 * t;
    ep_s[nt] = (MODE == TN_NN && d > 1) ? c / BNy : 0;
    ep_y[nt] = (MODE == TN_NN && d > 1) ? c % BNy : c;
  }

  long long w, w_end;
  if (p.split) {
    w = (long long)blockIdx.x * p.work_per_cta;
    w_end = min(w + p.work_per_cta, p.total_work);
  } else {
    w = blockIdx.x;
    w_end = p.total_tiles;
  }
  const long long w_step = p.split ? 0 : (long long)gridDim.x;  // split: segments follow each other; else tile stride

  // A work segment = consecutive k-blocks of one output tile.  Producer and consumers walk the same list of segments with
  // separate cursors: the producer (thread 0) stays STAGES-1 k-blocks ahead ACROSS segment boundaries, so a new tile never
  // starts with an empty pipeline.
  struct Segment {
    long long w;     // position in the work list
    int prob;        // problem index (monotone)
    int m0, n0;      // tile origin
    int link, k;     // next link / k-block inside the link
    int left;        // k-blocks left in this segment
    int n_it;        // k-blocks of the segment
  };
  auto decode = [&](Segment& sg) {  // sg.w, sg.prob given; fills the rest; returns false past the end
    if (sg.w >= w_end) return false;
    if (p.split) {
      while (sg.prob + 1 < p.n_problems && p.problems[sg.prob + 1].work_begin <= sg.w) ++sg.prob;
    } else {
      while (sg.prob + 1 < p.n_problems && p.problems[sg.prob + 1].tile_begin <= sg.w) ++sg.prob;
    }
    const ProblemDev& P = p.problems[sg.prob];
    const int iters_tile = P.link_count * p.ipl;
    int tile, i0;
    if (p.split) {
      const long long rem = sg.w - P.work_begin;
      tile = (int)(rem / iters_tile);
      i0 = (int)(rem % iters_tile);
      sg.n_it = (int)min((long long)(iters_tile - i0), w_end - sg.w);
    } else {
      tile = (int)(sg.w - P.tile_begin);
      i0 = 0;
      sg.n_it = iters_tile;
    }
    sg.m0 = (tile % p.tiles_m) * BMe;
    sg.n0 = (tile / p.tiles_m) * BNy;
    sg.link = P.link_begin + i0 / p.ipl;
    sg.k = i0 % p.ipl;
    sg.left = sg.n_it;
    return true;
  };
  auto next_segment = [&](Segment& sg) {
    sg.w += p.split ? sg.n_it : w_step;
    return decode(sg);
  };

  unsigned n_fill = 0, n_cons = 0;  // k-blocks filled / consumed by this CTA since kernel start (barrier phases, ring slots)
  Segment ps;                       // producer cursor (thread 0)
  ps.w = w; ps.prob = 0;
  bool p_valid = decode(ps);
  int li_cached = -1;
  const CUtensorMap *mapA = nullptr, *mapB = nullptr;
  const LinkDev* Lc = nullptr;

  // ---- producer: every thread tracks the cursor of the next k-block to stage (cheap integer state), so that ANY warp can
  // issue it.  A stage is refilled by the last consumer warp that finishes reading it (shared-memory arrival counter): the
  // load is issued at the earliest possible moment instead of when one fixed thread happens to come around. ----
  auto produce = [&](bool issuer) {
    if (!p_valid) return;
    if (ps.left == 0) {
      p_valid = next_segment(ps);
      if (!p_valid) return;
    }
    const int stage = n_fill % STAGES;
    if (issuer) {
      if (ps.link != li_cached) {
        Lc = p.links + ps.link;
        mapA = Lc->a_dyn ? &psi_map_a : maps + Lc->a_map;
        mapB = Lc->b_dyn ? &psi_map_b : maps + Lc->b_map;
        li_cached = ps.link;
      }
      {  // operator of this k-block's link, one ring slot per k-block in flight
        double* ring = sOpRing + (n_fill % kRing) * kOpSlot;
        const bool hop = (d > 1) && Lc->has_op;
        ring[kOpFlag] = hop ? 1.0 : 0.0;
        if (hop) {
#pragma unroll
          for (int i = 0; i < DMAX * DMAX; ++i) ring[i] = Lc->op[i];
        }
      }
      double* sA = smem + stage * STAGE_ELEMS;
      double* sB = sA + A_ELEMS;
      const int k0 = ps.k * BK;
      mbar_expect_tx(&full[stage], STAGE_BYTES);
#pragma unroll
      for (int h = 0; h < BK / BOXW; ++h) tma_2d(sA + h * BM * BOXW, mapA, k0 + h * BOXW, ps.m0, &full[stage]);
      if (MODE == TN_NT) {
#pragma unroll
        for (int h = 0; h < BK / BOXW; ++h) tma_2d(sB + h * BN * BOXW, mapB, k0 + h * BOXW, ps.n0, &full[stage]);
      } else {
        // B tile: BK k-rows, BN/16 boxes of 16 columns; box q covers tile columns [16q, 16q+16) = one (s, y) group
#pragma unroll
        for (int q = 0; q < BN / BOXW; ++q) {
          const int c = q * BOXW;
          const int sgrp = (d > 1) ? c / BNy : 0;
          const int y = (d > 1) ? c % BNy : c;
          if (sgrp < d)
            tma_3d(sB + q * BK * BOXW, mapB, ps.n0 + y, sgrp, k0, &full[stage]);
          else
            tma_3d(sB + q * BK * BOXW, mapB, Ny, 0, k0, &full[stage]);  // fully out of range: zero fill, keeps the byte count
        }
      }
    }
    if (++ps.k == p.ipl) {
      ps.k = 0;
      ++ps.link;
    }
    --ps.left;
    ++n_fill;
  };

  // the first STAGES k-blocks go into empty stages
#pragma unroll
  for (int s = 0; s < STAGES; ++s) produce(tid == 0);

  Segment cs;  // consumer cursor (all threads)
  cs.w = w; cs.prob = 0;
  bool c_valid = decode(cs);
  while (c_valid) {
    const ProblemDev P = p.problems[cs.prob];
    const int m0 = cs.m0, n0 = cs.n0, n_it = cs.n_it;

    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    // ---- consumer: one k4 slice ----
    auto compute_kk = [&](const double* sA, const double* sB, const double* sO, const bool has_op, const int kk) {
      const int h = kk >> 2, atom = kk & 3;  // 16-column box, 32-byte atom inside the 128-byte row
      double a[MT], b[NT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int r = warp_m * WM + mt * 8 + g;
        if (MODE == TN_NT && has_op) {
          double v = 0.0;
#pragma unroll
          for (int sp = 0; sp < DMAX; ++sp)
            if (sp < d) {
              const int rr = m_base[mt] + sp;
              v += sO[m_s[mt] * d + sp] * sA[h * (BM * BOXW) + rr * BOXW + ((atom ^ (rr & 3)) << 2) + t];
            }
          a[mt] = v;
        } else {
          a[mt] = sA[h * (BM * BOXW) + r * BOXW + ((atom ^ (g & 3)) << 2) + t];
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = warp_n * WN + nt * 8 + g;
        if (MODE == TN_NT) {
          b[nt] = sB[h * (BN * BOXW) + c * BOXW + ((atom ^ (g & 3)) << 2) + t];
        } else {
          const int k = kk * 4 + t;  // row of the KN tile, k & 3 == t
          if (has_op) {
            double v = 0.0;
#pragma unroll
            for (int sp = 0; sp < DMAX; ++sp)
              if (sp < d) {
                const int cc = sp * BNy + n_y[nt];
                v += sO[n_s[nt] * d + sp] * sB[(cc >> 4) * (BK * BOXW) + k * BOXW + ((((cc & 15) >> 2) ^ t) << 2) + (cc & 3)];
              }
            b[nt] = v;
          } else {
            b[nt] = sB[(c >> 4) * (BK * BOXW) + k * BOXW + ((((c & 15) >> 2) ^ t) << 2) + (c & 3)];
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt], a[mt], b[nt]);
    };

    for (int jj = 0; jj < n_it; ++jj) {
      const int stage = n_cons % STAGES;
      mbar_wait(&full[stage], (n_cons / STAGES) & 1);
      const double* sA = smem + stage * STAGE_ELEMS;
      const double* sB = sA + A_ELEMS;
      const double* sO = sOpRing + (n_cons % kRing) * kOpSlot;
      const bool has_op = sO[kOpFlag] != 0.0;
      if (has_op) {
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) compute_kk(sA, sB, sO, true, kk);
      } else {
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) compute_kk(sA, sB, sO, false, kk);
      }
      // this warp is done with the stage; the last of the THREADS/32 warps refills it with k-block n_cons + STAGES
      __syncwarp();
      int last = 0;
      if (lane == 0) {
        __threadfence_block();
        last = (atomicAdd(&done[2 * stage], 1) == THREADS / 32 - 1);
        if (last) {
          done[2 * stage] = 0;
          __threadfence_block();
        }
      }
      produce(last != 0);  // all threads advance the cursor; only the elected lane issues
      ++n_cons;
    }

    // ---- epilogue ----
    {
      double* Cg = P.c_dyn ? p.dyn_out : P.C;
      const double alpha = P.c_dyn ? P.alpha * p.dyn_alpha : P.alpha;
      const bool atomic = p.split != 0 || P.shared_out != 0;
      const bool vec_ok = !atomic && !P.accumulate && ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cg) & 15) == 0);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int r = warp_m * WM + mt * 8 + g;
        const int gm = m0 + r;
        if (r >= BMe || gm >= p.M) continue;
        double* crow = Cg + (size_t)gm * p.ldc;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          // columns 2t, 2t+1 of the 8-wide fragment: same (s) group, consecutive in memory
          int gn, nvalid;
          if (MODE == TN_NN && d > 1) {
            const int y = n0 + ep_y[nt];
            gn = ep_s[nt] * Ny + y;
            nvalid = ep_s[nt] < d ? min(max(Ny - y, 0), 2) : 0;
          } else {
            gn = n0 + warp_n * WN + nt * 8 + 2 * t;
            nvalid = min(max(p.N - gn, 0), 2);
          }
          if (nvalid == 0) continue;
          const double v0 = alpha * acc[mt][nt][0], v1 = alpha * acc[mt][nt][1];
          double* dst = crow + gn;
          if (vec_ok && nvalid == 2 && ((gn & 1) == 0)) {
            *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
          } else if (atomic) {
            atomicAdd(dst, v0);
            if (nvalid == 2) atomicAdd(dst + 1, v1);
          } else if (P.accumulate) {
            dst[0] += v0;
            if (nvalid == 2) dst[1] += v1;
          } else {
            dst[0] = v0;
            if (nvalid == 2) dst[1] = v1;
          }
        }
      }
    }
    c_valid = next_segment(cs);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(ptr);
  }
  return fn;
}

bool tma_available() { return encode_fn() != nullptr; }

// rows x cols row-major matrix (cols contiguous, leading dimension ld); box = 16 cols x box_rows
int tma_encode_2d(TmaMap* out, const double* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeFn fn = encode_fn();
  if (!fn) return TN_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)BOXW, (cuuint32_t)box_rows}, estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d, %lld x %lld, ld %lld) failed with %d", rows, cols, ld, (int)r);
    return TN_ERR_CUDA;
  }
  return TN_OK;
}

// (k, s, y) tensor: y contiguous (Ny), s stride Ny, k stride ld; box = 16 y x 1 s x BK k
int tma_encode_3d(TmaMap* out, const double* base, long long K, int d, long long Ny, long long ld) {
  EncodeFn fn = encode_fn();
  if (!fn) return TN_ERR_CUDA;
  cuuint64_t gdim[3] = {(cuuint64_t)Ny, (cuuint64_t)d, (cuuint64_t)K};
  cuuint64_t gstr[2] = {(cuuint64_t)Ny * sizeof(double), (cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[3] = {(cuuint32_t)BOXW, 1, (cuuint32_t)BK}, estr[3] = {1, 1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d, K %lld d %d Ny %lld ld %lld) failed with %d", K, d, Ny, ld, (int)r);
    return TN_ERR_CUDA;
  }
  return TN_OK;
}

int tma_box_rows_a() { return BM; }
int tma_box_rows_b_nt() { return BN; }

int gemm_launch_tma(const GemmLaunch& L, const GemmSchedule& S, const ProblemDev* problems_dev, const LinkDev* links_dev,
                    const TmaMap* maps_dev, const TmaMap& psi_a, const TmaMap& psi_b, const double* dyn_in, double* dyn_out,
                    double dyn_alpha, cudaStream_t stream) {
  GemmParams p;
  p.M = L.M; p.N = L.N; p.K = L.K; p.d = L.d;
  p.lda = L.lda; p.ldb = L.ldb; p.ldc = L.ldc;
  p.n_problems = L.n_problems;
  p.tiles_m = S.tiles_m; p.tiles_n = S.tiles_n; p.ipl = S.ipl;
  p.split = S.split;
  p.problems = problems_dev; p.links = links_dev;
  p.total_work = S.total_work; p.work_per_cta = S.work_per_cta; p.total_tiles = S.total_tiles;
  p.dyn_in = dyn_in; p.dyn_out = dyn_out; p.dyn_alpha = dyn_alpha;
  static bool configured = false;
  if (!configured) {
    TN_CUDA(cudaFuncSetAttribute(chain_gemm_tma_kernel<TN_NN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    TN_CUDA(cudaFuncSetAttribute(chain_gemm_tma_kernel<TN_NT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    TN_CUDA(cudaFuncSetAttribute(chain_gemm_tma_kernel<TN_NN, kMaxD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    TN_CUDA(cudaFuncSetAttribute(chain_gemm_tma_kernel<TN_NT, kMaxD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    configured = true;
  }
  const CUtensorMap* md = reinterpret_cast<const CUtensorMap*>(maps_dev);
  const CUtensorMap& pa = reinterpret_cast<const CUtensorMap&>(psi_a);
  const CUtensorMap& pb = reinterpret_cast<const CUtensorMap&>(psi_b);
  // the operator loops are unrolled to the physical dimension bound: d <= 2 (spin-1/2, the headline case) keeps the lean
  // instantiation, d = 3 (spin-1) and d = 4 (two-site window of spin-1/2) use the wide one
  const bool wide = L.d > 2;
  if (L.mode == TN_NN) {
    if (wide) chain_gemm_tma_kernel<TN_NN, kMaxD><<<S.grid, THREADS, SMEM_BYTES, stream>>>(p, md, pa, pb);
    else chain_gemm_tma_kernel<TN_NN, 2><<<S.grid, THREADS, SMEM_BYTES, stream>>>(p, md, pa, pb);
  } else {
    if (wide) chain_gemm_tma_kernel<TN_NT, kMaxD><<<S.grid, THREADS, SMEM_BYTES, stream>>>(p, md, pa, pb);
    else chain_gemm_tma_kernel<TN_NT, 2><<<S.grid, THREADS, SMEM_BYTES, stream>>>(p, md, pa, pb);
  }
  TN_LAUNCHED();
  return TN_OK;
}

}  // namespace tn
