// Chain GEMM on the FP64 tensor pipe of sm_100a.
//
//   C_p (+)= alpha_p * sum_{l in links(p)} opA_l(A_l) . opB_l(B_l)
//
// This single kernel family executes every dense contraction of the DMRG hot path: the two stages of the
// effective-Hamiltonian matvec (a1, MPSClass.py:755-776), the environment transfers (a5,
// TensorBasicModule.py:530-619) and the mode products (a6, TensorBasicModule.py:387-424).
//
// Hardware mapping (B200, measured in profiles/r01_fp64_peaks.txt): every f64 mma shape lowers to DMMA.8x8x4,
// issue interval 16 cycles per SM sub-partition, 37.1 TFLOP/s chip peak.  The kernel is therefore bound by DMMA
// issue, not by shared memory or L2: a CTA tile of 128x128x16 needs 32 KB of operands per 4096 DMMA cycles.
//   * operands are staged global -> shared with a 4-stage cp.async pipeline (16-byte copies, zero fill at the
//     ragged edges; an 8-byte variant serves odd leading dimensions such as the d^n bonds at the chain ends),
//   * fragments are read with conflict-free LDS.64 (row pitch = tile + 4 doubles),
//   * the d x d site operator of a link is applied while the fragment is built (the "physical-index reshape" of
//     the reference, absorb_matrix2tensor(..., 1), never touches memory),
//   * the sum over the links of a chain (coupling terms) runs inside the K loop of one output tile,
//   * work is distributed either as whole tiles (deterministic) or stream-K over (tile, link, k) iterations with
//     FP64 reductions (red.global.add.f64) so that 148 SMs stay busy when a launch has fewer tiles than SMs.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace tn {

// tile configuration of the large path; the defaults are the best of the variants measured on B200
// (profiles/r01_kernel_variants.md): 128x64 CTA tile, BK = 32, 2 stages, two independent CTAs per SM (ping-pong)
constexpr int BK = kBK;
constexpr int PAD = 4;

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// E doubles per copy: 2 -> 16-byte cp.async.cg, 1 -> 8-byte cp.async.ca.  src_bytes < size zero-fills the rest.
template <int E>
__device__ __forceinline__ void cp_async(void* smem, const void* gmem, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  if (E == 2)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int MINB_ = 1>
struct TileCfg {
  static constexpr int MINB = MINB_;
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int MT = WM / 8, NT = WN / 8;
};

template <class Cfg, int MODE>
struct SmemLayout {
  // rows x pitch of the A and B tiles of one stage
  static constexpr int A_ROWS = (MODE == TN_TN) ? BK : Cfg::BM;
  static constexpr int A_COLS = (MODE == TN_TN) ? Cfg::BM : BK;
  static constexpr int B_ROWS = (MODE == TN_NT) ? Cfg::BN : BK;
  static constexpr int B_COLS = (MODE == TN_NT) ? BK : Cfg::BN;
  static constexpr int A_PITCH = A_COLS + PAD, B_PITCH = B_COLS + PAD;
  static constexpr int A_ELEMS = A_ROWS * A_PITCH, B_ELEMS = B_ROWS * B_PITCH;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static constexpr int OP_ELEMS = 4 * kOpSlot;  // ring of 4 link operators
  static constexpr size_t BYTES = (size_t(STAGE_ELEMS) * Cfg::STAGES + OP_ELEMS) * sizeof(double);
};

// Tile copies.  Every thread owns one column group and PASSES rows, so all addressing is a per-thread constant plus a
// multiple of the row stride: no branches, out-of-range pieces become zero-byte (zero-filling) copies.
//   "MK": tile rows = M (or N) index, tile columns = K (contiguous in memory)
template <int ROWS, int PITCH, int THREADS, int E>
__device__ __forceinline__ void load_mk(double* s, const double* g, long long ld, int rows_valid, int kvalid, int tid,
                                        const double* dummy) {
  constexpr int CPR = BK / E, RSTEP = THREADS / CPR, PASSES = ROWS / RSTEP;
  static_assert(THREADS % CPR == 0 && ROWS % RSTEP == 0, "tile/thread mismatch");
  const int col = (tid % CPR) * E, r0 = tid / CPR;
  const int kbytes = min(max(kvalid - col, 0), E) * 8;
#pragma unroll
  for (int i = 0; i < PASSES; ++i) {
    const int r = r0 + i * RSTEP;
    const int bytes = r < rows_valid ? kbytes : 0;
    cp_async<E>(s + r * PITCH + col, bytes ? g + (long long)r * ld + col : dummy, bytes);
  }
}
//   "KN": tile rows = K, tile columns = N (or M) index (contiguous); gcol / cbytes are the per-thread global column
//   offset and valid byte count of its column group (they encode the (s,y) grouping of the NN operator mode)
template <int COLS, int PITCH, int THREADS, int E>
__device__ __forceinline__ void load_kn(double* s, const double* g, long long ld, int kvalid, int gcol, int cbytes, int tid,
                                        const double* dummy) {
  constexpr int CPR = COLS / E, RSTEP = THREADS / CPR, PASSES = BK / RSTEP;
  static_assert(THREADS % CPR == 0 && BK % RSTEP == 0, "tile/thread mismatch");
  const int col = (tid % CPR) * E, r0 = tid / CPR;
#pragma unroll
  for (int i = 0; i < PASSES; ++i) {
    const int r = r0 + i * RSTEP;
    const int bytes = r < kvalid ? cbytes : 0;
    cp_async<E>(s + r * PITCH + col, bytes ? g + (long long)r * ld + gcol : dummy, bytes);
  }
}

template <class Cfg, int MODE, bool A16, int DMAX>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) chain_gemm_kernel(const GemmParams p) {
  using SL = SmemLayout<Cfg, MODE>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, WM = Cfg::WM, WN = Cfg::WN, STAGES = Cfg::STAGES;
  constexpr int MT = Cfg::MT, NT = Cfg::NT, THREADS = Cfg::THREADS;
  constexpr int E = A16 ? 2 : 1;
  extern __shared__ __align__(16) double smem[];
  double* const sOpRing = smem + SL::STAGE_ELEMS * STAGES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int warp_m = warp % Cfg::WARPS_M, warp_n = warp / Cfg::WARPS_M;
  const int d = p.d;
  // effective tile extents when a d x d operator groups rows (NT) or columns (NN)
  const int BNy = (MODE == TN_NN && d > 1) ? ((BN / d) / 8) * 8 : BN;  // y values per tile (NN); N = d*Ny
  const int BMe = (MODE == TN_NT && d > 1) ? (BM / d) * d : BM;        // rows per tile (NT)
  const int Ny = (MODE == TN_NN && d > 1) ? p.N / d : p.N;
  const double* const dummy = reinterpret_cast<const double*>(p.links);  // valid global address for zero-byte copies

  // per-thread constants of the fragment <-> tile mapping
  int m_s[MT], m_base[MT];  // NT: physical index s of fragment row, first row of its (x, :) group
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    int r = warp_m * WM + mt * 8 + g;
    m_s[mt] = (MODE == TN_NT && d > 1) ? r % d : 0;
    m_base[mt] = r - m_s[mt];
  }
  int n_s[NT], n_y[NT];  // NN: physical index s of fragment column, y offset inside the tile
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    int c = warp_n * WN + nt * 8 + g;
    int s = (MODE == TN_NN && d > 1) ? c / BNy : 0;
    n_y[nt] = (MODE == TN_NN && d > 1) ? c % BNy : c;
    n_s[nt] = s < d ? s : 0;  // columns beyond d*BNy hold zeros and their results are discarded
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
    w = blockIdx.x;  // tile index
    w_end = p.total_tiles;
  }
  int prob = 0;

  while (w < w_end) {
    // ---- locate the work segment: problem, tile, iteration range ----
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
    const int n0 = tnn * BNy;  // NN with d>1: first y of the tile; otherwise first column

    // per-thread column group of the "KN" tiles (constant over the k loop)
    int a_gcol = 0, a_cbytes = 0, b_gcol = 0, b_cbytes = 0;
    if (MODE == TN_TN) {
      const int col = (tid % (BM / E)) * E;
      a_gcol = m0 + col;
      a_cbytes = min(max(p.M - m0 - col, 0), E) * 8;
    }
    if (MODE != TN_NT) {
      const int col = (tid % (BN / E)) * E;
      if (MODE == TN_NN && d > 1) {
        const int s = col / BNy, y = col % BNy;
        b_gcol = s * Ny + n0 + y;
        b_cbytes = (s < d) ? min(max(min(Ny - n0, BNy) - y, 0), E) * 8 : 0;
      } else {
        b_gcol = n0 + col;
        b_cbytes = min(max(p.N - n0 - col, 0), E) * 8;
      }
    }
    const int a_rows_valid = min(BMe, p.M - m0);  // MK tiles of A
    const int b_rows_valid = min(BN, p.N - n0);   // MK tile of B (NT)

    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    // ---- per-thread copy geometry (pass 0); pass i adds i * RSTEP rows ----
    constexpr int A_CPR = SL::A_COLS / E, A_RSTEP = THREADS / A_CPR, A_PASSES = SL::A_ROWS / A_RSTEP;
    constexpr int B_CPR = SL::B_COLS / E, B_RSTEP = THREADS / B_CPR, B_PASSES = SL::B_ROWS / B_RSTEP;
    const int a_r0 = tid / A_CPR, a_col = (tid % A_CPR) * E;
    const int b_r0 = tid / B_CPR, b_col = (tid % B_CPR) * E;
    // element offset of this thread's first copy relative to the operand base at k-block 0, advance per k-block
    const long long a_off0 = (MODE == TN_TN) ? (long long)a_r0 * p.lda + a_gcol : (long long)(m0 + a_r0) * p.lda + a_col;
    const long long b_off0 = (MODE == TN_NT) ? (long long)(n0 + b_r0) * p.ldb + b_col : (long long)b_r0 * p.ldb + b_gcol;
    const long long a_kadv = (MODE == TN_TN) ? (long long)BK * p.lda : BK;
    const long long b_kadv = (MODE == TN_NT) ? BK : (long long)BK * p.ldb;
    const long long a_pstride = (long long)A_RSTEP * p.lda, b_pstride = (long long)B_RSTEP * p.ldb;
    const int a_soff = a_r0 * SL::A_PITCH + a_col, b_soff = b_r0 * SL::B_PITCH + b_col;
    // interior tile: every copy of every k-block is a full, in-range 8*E-byte copy -> no predicates, pointer increments only
    bool fast = (p.K % BK == 0);
    if (MODE == TN_TN) fast = fast && (p.M - m0 >= BM); else fast = fast && (a_rows_valid == BM);
    if (MODE == TN_NT) fast = fast && (b_rows_valid == BN);
    else if (MODE == TN_NN && d > 1) fast = fast && (d * BNy == BN) && (Ny - n0 >= BNy);
    else fast = fast && (p.N - n0 >= BN);

    // ---- producer / consumer positions inside the link chain, advanced incrementally ----
    int p_link = P.link_begin + i0 / p.ipl, p_k = i0 % p.ipl, p_left = n_it;
    int c_k = p_k, c_link = p_link;
    int li_cached = -1;
    const double *Ag = nullptr, *Bg = nullptr;  // operand bases of the cached link
    const double *pa = nullptr, *pb = nullptr;  // fast path: this thread's pass-0 source pointers for the next k-block

    auto issue = [&](int stage) {
      if (p_left > 0) {  // CTA-uniform
        if (p_link != li_cached) {  // rare: first k-block of a link
          const LinkDev* L = p.links + p_link;
          Ag = L->a_dyn ? p.dyn_in : L->A;
          Bg = L->b_dyn ? p.dyn_in : L->B;
          if (tid < kOpSlot) sOpRing[(p_link & 3) * kOpSlot + tid] = tid < kMaxD * kMaxD ? L->op[tid] : (tid == kOpFlag ? (double)L->has_op : 0.0);
          li_cached = p_link;
          pa = Ag + a_off0 + p_k * a_kadv;
          pb = Bg + b_off0 + p_k * b_kadv;
        }
        double* sA = smem + stage * SL::STAGE_ELEMS;
        double* sB = sA + SL::A_ELEMS;
        if (fast) {
          const double* q = pa;
#pragma unroll
          for (int i = 0; i < A_PASSES; ++i) {
            cp_async<E>(sA + a_soff + i * A_RSTEP * SL::A_PITCH, q, E * 8);
            q += a_pstride;
          }
          q = pb;
#pragma unroll
          for (int i = 0; i < B_PASSES; ++i) {
            cp_async<E>(sB + b_soff + i * B_RSTEP * SL::B_PITCH, q, E * 8);
            q += b_pstride;
          }
          pa += a_kadv;
          pb += b_kadv;
        } else {
          const int k0 = p_k * BK;
          const int kvalid = p.K - k0;
          if (MODE == TN_TN)
            load_kn<BM, SL::A_PITCH, THREADS, E>(sA, Ag + (long long)k0 * p.lda, p.lda, kvalid, a_gcol, a_cbytes, tid, dummy);
          else
            load_mk<BM, SL::A_PITCH, THREADS, E>(sA, Ag + (long long)m0 * p.lda + k0, p.lda, a_rows_valid, kvalid, tid, dummy);
          if (MODE == TN_NT)
            load_mk<BN, SL::B_PITCH, THREADS, E>(sB, Bg + (long long)n0 * p.ldb + k0, p.ldb, b_rows_valid, kvalid, tid, dummy);
          else
            load_kn<BN, SL::B_PITCH, THREADS, E>(sB, Bg + (long long)k0 * p.ldb, p.ldb, kvalid, b_gcol, b_cbytes, tid, dummy);
        }
        if (++p_k == p.ipl) {
          p_k = 0;
          ++p_link;
        }
        --p_left;
      }
      cp_async_commit();
    };

    // ---- consumer: DMMA over one k4 slice of a staged k-block ----
    auto compute_kk = [&](const double* sA, const double* sB, const double* sO, const bool has_op, const int kk) {
      const int kc = kk * 4 + t;
      double a[MT], b[NT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int r = warp_m * WM + mt * 8 + g;
        if (MODE == TN_TN) {
          a[mt] = sA[kc * SL::A_PITCH + r];
        } else if (MODE == TN_NT && has_op) {
          double v = 0.0;
#pragma unroll
          for (int sp = 0; sp < DMAX; ++sp)
            if (sp < d) v += sO[m_s[mt] * d + sp] * sA[(m_base[mt] + sp) * SL::A_PITCH + kc];
          a[mt] = v;
        } else {
          a[mt] = sA[r * SL::A_PITCH + kc];
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = warp_n * WN + nt * 8 + g;
        if (MODE == TN_NT) {
          b[nt] = sB[c * SL::B_PITCH + kc];
        } else if (MODE == TN_NN && has_op) {
          double v = 0.0;
#pragma unroll
          for (int sp = 0; sp < DMAX; ++sp)
            if (sp < d) v += sO[n_s[nt] * d + sp] * sB[kc * SL::B_PITCH + sp * BNy + n_y[nt]];
          b[nt] = v;
        } else {
          b[nt] = sB[kc * SL::B_PITCH + c];
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt], a[mt], b[nt]);
    };

    // ---- software pipeline: the copies of k-block jj+STAGES-1 are issued between the k4 slices of k-block jj ----
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    const double* sO = sOpRing;
    bool has_op = false;
    for (int jj = 0; jj < n_it; ++jj) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      const int stage = jj % STAGES;
      const double* sA = smem + stage * SL::STAGE_ELEMS;
      const double* sB = sA + SL::A_ELEMS;
      if (jj == 0 || c_k == 0) {  // first k-block of a link: pick up its operator
        sO = sOpRing + (c_link & 3) * kOpSlot;
        has_op = (d > 1) && (MODE != TN_TN) && (sO[kOpFlag] != 0.0);
      }
      if (++c_k == p.ipl) {
        c_k = 0;
        ++c_link;
      }
      const int nstage = (jj + STAGES - 1) % STAGES;
      if (has_op) {
        compute_kk(sA, sB, sO, true, 0);
        issue(nstage);
#pragma unroll
        for (int kk = 1; kk < BK / 4; ++kk) compute_kk(sA, sB, sO, true, kk);
      } else {
        compute_kk(sA, sB, sO, false, 0);
        issue(nstage);
#pragma unroll
        for (int kk = 1; kk < BK / 4; ++kk) compute_kk(sA, sB, sO, false, kk);
      }
    }
    cp_async_wait<0>();
    __syncthreads();

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
    w += p.split ? n_it : gridDim.x;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#ifndef TN_CFGL_WN
#define TN_CFGL_WN 64
#endif
using CfgL = TileCfg<TN_CFGL_BM, TN_CFGL_BN, 32, TN_CFGL_WN, TN_STAGES_L, TN_CFGL_CTAS>;  // default: 4 warps of 32x64, 64 accumulator doubles per thread
using CfgS = TileCfg<64, 64, 32, 32, (TN_BK > 16 ? 3 : 4)>;    // 128 threads, 32 accumulator doubles per thread

template <class Cfg, int MODE, bool A16, int DMAX>
static int launch_one(const GemmParams& p, int grid, cudaStream_t stream) {
  using SL = SmemLayout<Cfg, MODE>;
  static bool configured = false;
  auto kern = chain_gemm_kernel<Cfg, MODE, A16, DMAX>;
  if (!configured) {
    TN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::BYTES));
    configured = true;
  }
  kern<<<grid, Cfg::THREADS, SL::BYTES, stream>>>(p);
  TN_LAUNCHED();
  return TN_OK;
}

template <class Cfg, bool A16>
static int launch_mode(int mode, const GemmParams& p, int grid, cudaStream_t stream) {
  switch (mode) {
    // operator loops unrolled to 2 (spin-1/2) or kMaxD (spin-1, two-site window); TN carries no operator
    case TN_NN: return p.d > 2 ? launch_one<Cfg, TN_NN, A16, kMaxD>(p, grid, stream) : launch_one<Cfg, TN_NN, A16, 2>(p, grid, stream);
    case TN_NT: return p.d > 2 ? launch_one<Cfg, TN_NT, A16, kMaxD>(p, grid, stream) : launch_one<Cfg, TN_NT, A16, 2>(p, grid, stream);
    case TN_TN: return launch_one<Cfg, TN_TN, A16, 2>(p, grid, stream);
  }
  set_error("chain_gemm: unknown mode %d", mode);
  return TN_ERR_INVALID;
}

static void tile_counts(int mode, int M, int N, int d, int BM, int BN, int* tm, int* tn_) {
  int BMe = (mode == TN_NT && d > 1) ? (BM / d) * d : BM;
  *tm = (M + BMe - 1) / BMe;
  if (mode == TN_NN && d > 1) {
    int BNy = ((BN / d) / 8) * 8, Ny = N / d;
    *tn_ = (Ny + BNy - 1) / BNy;
  } else {
    *tn_ = (N + BN - 1) / BN;
  }
}

int gemm_plan_schedule(const GemmLaunch& L, ProblemDev* problems, const LinkDev* links, const double* dyn_in,
                       const double* dyn_out, GemmSchedule* S) {
  TN_REQUIRE(L.mode >= 0 && L.mode <= 2, "chain_gemm: mode %d", L.mode);
  TN_REQUIRE(L.M > 0 && L.N > 0 && L.K > 0, "chain_gemm: empty shape M=%d N=%d K=%d", L.M, L.N, L.K);
  TN_REQUIRE(L.d >= 1 && L.d <= kMaxD, "chain_gemm: physical dimension %d not in 1..%d", L.d, kMaxD);
  TN_REQUIRE(L.n_problems > 0 && L.n_links > 0, "chain_gemm: no work");
  if (L.mode == TN_NN && L.d > 1) TN_REQUIRE(L.N % L.d == 0, "chain_gemm NN: N=%d not a multiple of d=%d", L.N, L.d);
  if (L.mode == TN_NT && L.d > 1) TN_REQUIRE(L.M % L.d == 0, "chain_gemm NT: M=%d not a multiple of d=%d", L.M, L.d);
  const int minA = (L.mode == TN_TN) ? L.M : L.K;
  const int minB = (L.mode == TN_NT) ? L.K : L.N;
  TN_REQUIRE(L.lda >= minA && L.ldb >= minB && L.ldc >= L.N, "chain_gemm: leading dimension too small");

  // alignment: the 16-byte copy path needs even pitches, even group offsets and 16-byte aligned bases
  bool a16 = (L.lda % 2 == 0) && (L.ldb % 2 == 0);
  if (L.mode == TN_NN && L.d > 1) a16 = a16 && ((L.N / L.d) % 2 == 0);
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  for (int l = 0; l < L.n_links; ++l) {
    const double* A = links[l].a_dyn ? dyn_in : links[l].A;
    const double* B = links[l].b_dyn ? dyn_in : links[l].B;
    TN_REQUIRE(A && B, "chain_gemm: link %d has a null operand", l);
    a16 = a16 && al(A) && al(B);
  }
  long long total_links = 0;
  for (int q = 0; q < L.n_problems; ++q) {
    TN_REQUIRE(problems[q].link_count > 0 && problems[q].link_begin >= 0 &&
                   problems[q].link_begin + problems[q].link_count <= L.n_links,
               "chain_gemm: problem %d has a bad link range", q);
    TN_REQUIRE(problems[q].c_dyn ? dyn_out != nullptr : problems[q].C != nullptr, "chain_gemm: problem %d has no output", q);
    total_links += problems[q].link_count;
  }

  const int sms = sm_count();
  // tile configuration: large tiles when they still give every SM work, else small tiles
  int tmL, tnL, tmS, tnS;
  tile_counts(L.mode, L.M, L.N, L.d, CfgL::BM, CfgL::BN, &tmL, &tnL);
  tile_counts(L.mode, L.M, L.N, L.d, CfgS::BM, CfgS::BN, &tmS, &tnS);
  const int ipl = (L.K + BK - 1) / BK;
  const long long workL = (long long)tmL * tnL * total_links * ipl;
  // use the large configuration when each SM would get at least ~8 k-iterations of a large tile
  const bool large = (L.M >= 96 && L.N >= 96) && workL >= 8LL * sms;
  S->config = large ? 0 : 1;
  S->aligned16 = a16 ? 1 : 0;
  S->tiles_m = large ? tmL : tmS;
  S->tiles_n = large ? tnL : tnS;
  S->ipl = ipl;
  const int tiles = S->tiles_m * S->tiles_n;
  const int ctas_per_sm = large ? TN_CFGL_CTAS : 2;
  // launches that share the SMs with a concurrent one (grid_div) take their share of the CTA slots in the small-tile
  // configuration only: measured -8 % per Lanczos step at chi = 256, +3 % at chi = 512 (large tiles), profiles/r02_small_chi.md
  const int max_grid = std::max(1, sms * ctas_per_sm / (large ? 1 : std::max(1, L.grid_div)));

  long long work = 0, tile_acc = 0, max_tile_iters = 0;
  for (int q = 0; q < L.n_problems; ++q) {
    problems[q].work_begin = work;
    problems[q].tile_begin = tile_acc;
    long long it = (long long)problems[q].link_count * ipl;
    work += (long long)tiles * it;
    tile_acc += tiles;
    max_tile_iters = std::max(max_tile_iters, it);
  }
  S->total_work = work;
  S->total_tiles = tile_acc;

  // whole-tile schedule: tile w goes to CTA w % grid.  Estimate its makespan.
  int grid_tiles = (int)std::min<long long>(tile_acc, max_grid);
  std::vector<long long> load(grid_tiles, 0);
  {
    long long w = 0;
    for (int q = 0; q < L.n_problems; ++q) {
      long long it = (long long)problems[q].link_count * ipl;
      for (int tI = 0; tI < tiles; ++tI, ++w) load[w % grid_tiles] += it;
    }
  }
  long long makespan = *std::max_element(load.begin(), load.end());
  double eff_tiles = (double)work / ((double)makespan * max_grid);
  bool split = !L.deterministic && eff_tiles < 0.90 && work >= 2;
  if (split) {
    // stream-K: every CTA gets the same number of k-iterations (at least 4 to amortise the pipeline fill)
    long long per = std::max<long long>((work + max_grid - 1) / max_grid, 4);
    int grid = (int)((work + per - 1) / per);
    S->split = 1;
    S->grid = grid;
    S->work_per_cta = per;
  } else {
    S->split = 0;
    S->grid = grid_tiles;
    S->work_per_cta = 0;
  }
  return TN_OK;
}

int gemm_launch(const GemmLaunch& L, const GemmSchedule& S, const ProblemDev* problems_dev, const LinkDev* links_dev,
                const double* dyn_in, double* dyn_out, double dyn_alpha, cudaStream_t stream) {
  GemmParams p;
  p.M = L.M; p.N = L.N; p.K = L.K; p.d = L.d;
  p.lda = L.lda; p.ldb = L.ldb; p.ldc = L.ldc;
  p.n_problems = L.n_problems;
  p.tiles_m = S.tiles_m; p.tiles_n = S.tiles_n; p.ipl = S.ipl;
  p.split = S.split;
  p.problems = problems_dev; p.links = links_dev;
  p.total_work = S.total_work; p.work_per_cta = S.work_per_cta; p.total_tiles = S.total_tiles;
  p.dyn_in = dyn_in; p.dyn_out = dyn_out; p.dyn_alpha = dyn_alpha;
  if (S.config == 0)
    return S.aligned16 ? launch_mode<CfgL, true>(L.mode, p, S.grid, stream) : launch_mode<CfgL, false>(L.mode, p, S.grid, stream);
  return S.aligned16 ? launch_mode<CfgS, true>(L.mode, p, S.grid, stream) : launch_mode<CfgS, false>(L.mode, p, S.grid, stream);
}

}  // namespace tn

// ------------------------------------------------------------------------------------------------
// public C ABI: generic chain GEMM
// ------------------------------------------------------------------------------------------------
using namespace tn;

extern "C" size_t tn_chain_gemm_workspace_bytes(int n_problems, int n_links) {
  return align_up(sizeof(ProblemDev) * (size_t)std::max(n_problems, 1)) + align_up(sizeof(LinkDev) * (size_t)std::max(n_links, 1));
}

extern "C" int tn_chain_gemm(int mode, int M, int N, int K, int d, int lda, int ldb, int ldc, const tn_problem* problems,
                             int n_problems, const tn_link* links, int n_links, int deterministic, void* workspace,
                             size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(problems && links && n_problems > 0 && n_links > 0, "tn_chain_gemm: empty problem/link list");
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_chain_gemm: workspace must be 256-byte aligned");
  if (workspace_bytes < tn_chain_gemm_workspace_bytes(n_problems, n_links)) {
    set_error("tn_chain_gemm: workspace %zu < %zu bytes", workspace_bytes, tn_chain_gemm_workspace_bytes(n_problems, n_links));
    return TN_ERR_WORKSPACE;
  }
  std::vector<ProblemDev> ph(n_problems);
  std::vector<LinkDev> lh(n_links);
  for (int l = 0; l < n_links; ++l) {
    lh[l].A = links[l].A; lh[l].B = links[l].B;
    for (int i = 0; i < kMaxD * kMaxD; ++i) lh[l].op[i] = links[l].op[i];
    lh[l].has_op = links[l].has_op && d > 1;
    lh[l].a_dyn = lh[l].b_dyn = lh[l].pad = 0;
  }
  for (int q = 0; q < n_problems; ++q) {
    ph[q].C = problems[q].C; ph[q].alpha = problems[q].alpha;
    ph[q].link_begin = problems[q].link_begin; ph[q].link_count = problems[q].link_count;
    ph[q].accumulate = problems[q].accumulate; ph[q].c_dyn = 0; ph[q].shared_out = 0; ph[q].pad = 0;
  }
  GemmLaunch L{mode, M, N, K, d, lda, ldb, ldc, n_problems, n_links, deterministic};
  GemmSchedule S;
  TN_CHECK(gemm_plan_schedule(L, ph.data(), lh.data(), nullptr, nullptr, &S));
  if (S.split) {  // partial tiles are combined with atomics: outputs that are not accumulated start from zero
    for (int q = 0; q < n_problems; ++q)
      if (!ph[q].accumulate) TN_CUDA(cudaMemset2DAsync(ph[q].C, sizeof(double) * ldc, 0, sizeof(double) * N, M, stream));
  }
  Carver cw(workspace, workspace_bytes);
  ProblemDev* pd = cw.take<ProblemDev>(n_problems);
  LinkDev* ld = cw.take<LinkDev>(n_links);
  TN_CUDA(cudaMemcpyAsync(pd, ph.data(), sizeof(ProblemDev) * n_problems, cudaMemcpyHostToDevice, stream));
  TN_CUDA(cudaMemcpyAsync(ld, lh.data(), sizeof(LinkDev) * n_links, cudaMemcpyHostToDevice, stream));
  return gemm_launch(L, S, pd, ld, nullptr, nullptr, 1.0, stream);
}
