// TMA-staged variant of the chain GEMM (modes TN_NN and TN_NT) used by the effective-Hamiltonian matvec.
//
// Same math, tiling (128x64x32, 4 warps of 32x64, two CTAs per SM) and DMMA inner loop as chain_gemm.cu; the operand
// staging is done by the Tensor Memory Accelerator instead of per-thread cp.async:
//   * one elected thread issues 4-6 `cp.async.bulk.tensor` box loads (16 doubles = 128 B wide) per k-block; the boxes land
//     in shared memory in the SWIZZLE_128B_ATOM_32B layout: the 32-byte atom `a` of row `r` sits at position a ^ (r & 3)
//     (measured with tools/microbench/tma_probe.cu), which makes every LDS.64 fragment read conflict-free for both tile
//     orientations without padding;
//   * completion is tracked by `full`/`empty` mbarriers per stage: the consumer warps never execute a copy instruction and
//     never meet at a CTA-wide barrier inside the k loop;
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
constexpr int BM = 128, BN = 64, BK = 32, WM = 32, WN = 64, STAGES = 2, THREADS = 128;
constexpr int MT = WM / 8, NT = WN / 8;
constexpr int BOXW = 16;                      // doubles per box row (128 bytes)
constexpr int A_ELEMS = BM * BK;              // two boxes of BM x 16
constexpr int B_ELEMS = BN * BK;              // NT: two boxes of BN x 16; NN: four boxes of BK x 16
constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
constexpr unsigned STAGE_BYTES = STAGE_ELEMS * sizeof(double);
constexpr size_t SMEM_BYTES = size_t(STAGE_ELEMS) * STAGES * sizeof(double) + 4 * 16 * sizeof(double) + 2 * STAGES * sizeof(uint64_t);

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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) chain_gemm_tma_kernel(const GemmParams p, const CUtensorMap* __restrict__ maps,
                                                                    const __grid_constant__ CUtensorMap psi_map_a,
                                                                    const __grid_constant__ CUtensorMap psi_map_b) {
  extern __shared__ __align__(1024) double smem[];  // the swizzle pattern is a function of the shared address: 1024-byte aligned base
  if (smem_u32(smem) & 1023u) __trap();
  double* const sOpRing = smem + STAGE_ELEMS * STAGES;
  uint64_t* const full = reinterpret_cast<uint64_t*>(sOpRing + 4 * 16);
  uint64_t* const empty = full + STAGES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int warp_m = warp;  // 4 x 1 warps
  const int d = p.d;
  const int BNy = (MODE == TN_NN && d > 1) ? ((BN / d) / BOXW) * BOXW : BN;  // y values per tile (multiple of the box width)
  const int BMe = (MODE == TN_NT && d > 1) ? (BM / d) * d : BM;
  const int Ny = (MODE == TN_NN && d > 1) ? p.N / d : p.N;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], THREADS / 32);
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
    const int c = nt * 8 + g;
    const int s = (MODE == TN_NN && d > 1) ? c / BNy : 0;
    n_y[nt] = (MODE == TN_NN && d > 1) ? c % BNy : c;
    n_s[nt] = s < d ? s : 0;
  }

  long long w, w_end;
  if (p.split) {
    w = (long long)blockIdx.x * p.work_per_cta;
    w_end = min(w + p.work_per_cta, p.total_work);
  } else {
    w = blockIdx.x;
    w_end = p.total_tiles;
  }
  int prob = 0;
  unsigned n_fill = 0, n_cons = 0;  // k-blocks filled / consumed by this CTA since kernel start (barrier phases)

  while (w < w_end) {
    int tile, i0, n_it;
    if (p.split) {
      while (prob + 1 < p.n_problems && p.problems[prob + 1].work_begin <= w) ++prob;
    } else {
      while (prob + 1 < p.n_problems && p.problems[prob + 1].tile_begin <= w) ++prob;
    }
    const ProblemDev P = p.problems[prob];
    const int iters_tile = P.link_count * p.ipl;
    if (p.split) {
      long long rem = w - P.work_begin;
      tile = (int)(rem / iters_tile);
      i0 = (int)(rem % iters_tile);
      n_it = (int)min((long long)(iters_tile - i0), w_end - w);
    } else {
      tile = (int)(w - P.tile_begin);
      i0 = 0;
      n_it = iters_tile;
    }
    const int tm = tile % p.tiles_m, tnn = tile / p.tiles_m;
    const int m0 = tm * BMe;
    const int n0 = tnn * BNy;

    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    int p_link = P.link_begin + i0 / p.ipl, p_k = i0 % p.ipl, p_left = n_it;
    int c_k = p_k, c_link = p_link;
    int li_cached = -1;
    const CUtensorMap *mapA = nullptr, *mapB = nullptr;

    // ---- producer (thread 0 only) ----
    auto produce = [&]() {
      if (p_left <= 0) return;
      const int stage = n_fill % STAGES;
      if (n_fill >= STAGES) mbar_wait(&empty[stage], ((n_fill / STAGES) - 1) & 1);
      if (p_link != li_cached) {
        const LinkDev* L = p.links + p_link;
        mapA = L->a_dyn ? &psi_map_a : maps + L->a_map;
        mapB = L->b_dyn ? &psi_map_b : maps + L->b_map;
        double* ring = sOpRing + (p_link & 3) * 16;
#pragma unroll
        for (int i = 0; i < kMaxD * kMaxD; ++i) ring[i] = L->op[i];
        ring[15] = (double)L->has_op;
        li_cached = p_link;
      }
      double* sA = smem + stage * STAGE_ELEMS;
      double* sB = sA + A_ELEMS;
      const int k0 = p_k * BK;
      mbar_expect_tx(&full[stage], STAGE_BYTES);
      // A tile: rows m0 .. m0+BM, two boxes of 16 k-columns
      tma_2d(sA, mapA, k0, m0, &full[stage]);
      tma_2d(sA + BM * BOXW, mapA, k0 + BOXW, m0, &full[stage]);
      if (MODE == TN_NT) {
        tma_2d(sB, mapB, k0, n0, &full[stage]);
        tma_2d(sB + BN * BOXW, mapB, k0 + BOXW, n0, &full[stage]);
      } else {
        // B tile: BK k-rows, BN/16 boxes of 16 columns; box q covers tile columns [16q, 16q+16) = (s, y) group
#pragma unroll
        for (int q = 0; q < BN / BOXW; ++q) {
          const int c = q * BOXW;
          const int s = (d > 1) ? c / BNy : 0;
          const int y = (d > 1) ? c % BNy : c;
          if (s < d)
            tma_3d(sB + q * BK * BOXW, mapB, n0 + y, s, k0, &full[stage]);
          else
            tma_3d(sB + q * BK * BOXW, mapB, Ny, 0, k0, &full[stage]);  // fully out of range: zero fill, keeps the byte count
        }
      }
      if (++p_k == p.ipl) {
        p_k = 0;
        ++p_link;
      }
      --p_left;
      ++n_fill;
    };

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
          for (int sp = 0; sp < kMaxD; ++sp)
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
        const int c = nt * 8 + g;
        if (MODE == TN_NT) {
          b[nt] = sB[h * (BN * BOXW) + c * BOXW + ((atom ^ (g & 3)) << 2) + t];
        } else {
          const int k = kk * 4 + t;  // row of the KN tile, k & 3 == t
          if (has_op) {
            double v = 0.0;
#pragma unroll
            for (int sp = 0; sp < kMaxD; ++sp)
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

    // ---- pipeline ----
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < STAGES - 1; ++s) produce();
    }
    const double* sO = sOpRing;
    bool has_op = false;
    for (int jj = 0; jj < n_it; ++jj) {
      if (tid == 0) produce();  // k-block jj + STAGES - 1
      const int stage = n_cons % STAGES;
      mbar_wait(&full[stage], (n_cons / STAGES) & 1);
      const double* sA = smem + stage * STAGE_ELEMS;
      const double* sB = sA + A_ELEMS;
      if (jj == 0 || c_k == 0) {
        sO = sOpRing + (c_link & 3) * 16;
        has_op = (d > 1) && (sO[15] != 0.0);
      }
      if (++c_k == p.ipl) {
        c_k = 0;
        ++c_link;
      }
      if (has_op) {
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) compute_kk(sA, sB, sO, true, kk);
      } else {
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) compute_kk(sA, sB, sO, false, kk);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      ++n_cons;
    }

    // ---- epilogue ----
    double* Cg = P.c_dyn ? p.dyn_out : P.C;
    const double alpha = P.c_dyn ? P.alpha * p.dyn_alpha : P.alpha;
    const bool atomic = p.split != 0 || P.shared_out != 0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r = warp_m * WM + mt * 8 + g;
      const int gm = m0 + r;
      if (r >= BMe || gm >= p.M) continue;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = nt * 8 + 2 * t + e;
          int gn;
          if (MODE == TN_NN && d > 1) {
            const int s = c / BNy, y = n0 + c % BNy;
            if (s >= d || y >= Ny) continue;
            gn = s * Ny + y;
          } else {
            gn = n0 + c;
            if (gn >= p.N) continue;
          }
          double* dst = Cg + (size_t)gm * p.ldc + gn;
          const double v = alpha * acc[mt][nt][e];
          if (atomic)
            atomicAdd(dst, v);
          else if (P.accumulate)
            *dst += v;
          else
            *dst = v;
        }
      }
    }
    w += p.split ? n_it : gridDim.x;
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
    TN_CUDA(cudaFuncSetAttribute(chain_gemm_tma_kernel<TN_NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    TN_CUDA(cudaFuncSetAttribute(chain_gemm_tma_kernel<TN_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    configured = true;
  }
  const CUtensorMap* md = reinterpret_cast<const CUtensorMap*>(maps_dev);
  const CUtensorMap& pa = reinterpret_cast<const CUtensorMap&>(psi_a);
  const CUtensorMap& pb = reinterpret_cast<const CUtensorMap&>(psi_b);
  if (L.mode == TN_NN)
    chain_gemm_tma_kernel<TN_NN><<<S.grid, THREADS, SMEM_BYTES, stream>>>(p, md, pa, pb);
  else
    chain_gemm_tma_kernel<TN_NT><<<S.grid, THREADS, SMEM_BYTES, stream>>>(p, md, pa, pb);
  TN_LAUNCHED();
  return TN_OK;
}

}  // namespace tn
