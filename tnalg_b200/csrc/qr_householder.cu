// Blocked Householder QR of the gauge moves (a7: left2right/right2left_decompose_tensor with way='qr',
// TensorBasicModule.py:342-345,378-380 -> np.linalg.qr; every sweep step of DMRG_anyH.py:47-64 calls it).
//
//   A (m, n) = Q (m, k) R (k, n),  k = min(m, n),  Q^T Q = 1 to rounding for ANY conditioning (centre tensors carry the
//   Schmidt spectrum, kappa up to 1e12 and beyond: Gram-matrix / Cholesky variants are not acceptable here).
//   Same reflector convention as LAPACK dgeqr2/dlarfg (beta = -sign(alpha) |x|), so R agrees with np.linalg.qr including
//   the signs of its rows whenever A has full column rank.
//
// Structure (right-looking, panel width 32):
//   qr_panel_kernel   one thread-block CLUSTER per panel (128 rows per CTA up to 1024 rows, 256 at 2048).  The panel lives in
//                     REGISTERS (thread = (row slot, column)).  Per column ONE reduction -- warp partials through shared memory,
//                     CTA partials through asynchronous remote stores that count on the destination's mbarrier -- yields the
//                     whole row  g_k = sum_{r>=j} P[r,j] P[r,k]  of the panel Gram matrix, from which the column norm (k = j), the
//                     reflector products v^T a_k (k > j) and the entries V_k^T v_j of the compact-WY factor T (k < j) all follow:
//                     y_k = (g_k - beta P[j,k]) / (alpha - beta).  Before factoring, the kernel applies the block reflector of the
//                     PREVIOUS panel to its own columns in registers (fused look-ahead), so no launch sits between two panels.
//   qr_apply_kernel   C <- (1 - V op(T) V^T) C on a strip of 32 columns per cluster (the CTAs split the rows) with FP64
//                     tensor-core MMAs (DMMA.8x8x4): W = V^T C (partial tiles meet in distributed shared memory, summed in
//                     CTA order), W <- op(T) W, C -= V W.  Carries the wide trailing update (op(T) = T^T) on a side stream while
//                     the next panel is factored.
//   form_q_compact_wy Q = [1; 0] - V (T V1^T) with T^-1 = diag(1 / tau) + striu(V^T V): deterministic chain GEMMs plus a
//                     recursive-doubling inverse over the 32 x 32 diagonal blocks (short matrices form Q by reverse applies).
// No atomics anywhere: results are bit-reproducible (replicas on several GPUs stay identical).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "vector_ops.cuh"

namespace cg = cooperative_groups;

namespace tn {

namespace {
constexpr int QNB = 32;        // panel width
constexpr int QMAXC = 8;       // portable cluster size
constexpr int QSLOT = 2 * QNB; // doubles one CTA contributes per column: g[32], row j[32]
constexpr int QMAX_RPT = 48;   // rows per thread of the panel kernel -> 768 rows per CTA, 6144 per cluster
constexpr int QNC = 32;        // strip width of the apply kernel
constexpr int ATHREADS = 256;  // apply kernel: 8 warps
constexpr int AWARPS = ATHREADS / 32;
constexpr int AVP = QNB + 4;   // pitch of the reflector block in shared memory (conflict-free fragment reads)

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// cluster-wide barrier with release / acquire ordering at CLUSTER scope.  cooperative_groups' cluster_barrier() emits a GPU-scope
// fence (MEMBAR.ALL.GPU) in front of the hardware barrier, which dominated the per-column exchange of the panel kernel.
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// ---- asynchronous cluster exchange: remote shared-memory stores that signal an mbarrier of the destination CTA (STAS) ----
__device__ __forceinline__ unsigned smem_u32q(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init_q(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32q(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_q(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32q(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_q(uint64_t* bar, unsigned parity) {
  unsigned done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32q(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) __trap();  // a lost signal must abort the kernel, not hang the GPU
  }
}
// 8-byte store into the shared memory of CTA `dst` of the cluster; completes 8 bytes on that CTA's mbarrier `bar`
__device__ __forceinline__ void st_async_q(double* local_addr, double v, uint64_t* bar, unsigned dst) {
  unsigned ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32q(local_addr)), "r"(dst));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32q(bar)), "r"(dst));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(ra), "d"(v), "r"(rb) : "memory");
}

}  // namespace

// W: work matrix (m x n, leading dimension ld); panel = columns [j0, j0 + nbp), rows [j0, m).
// Register-resident panel: thread (warp w, lane k) holds column k of the rows  row_lo + i*16 + w  (i < RPT) of this CTA, so the
// column-j broadcast is a warp shuffle and the rank-1 update touches registers only; shared memory carries just the
// cross-warp / cross-CTA partial sums (double-buffered by the parity of j: ONE CTA barrier -- plus one cluster barrier when
// the panel spans several CTAs -- per column).
#ifdef TN_QR_TIMING
__device__ long long g_qr_clk[8];
__device__ long long g_qr_aclk[8];   // apply kernel: phases 0..6, [7] = launches
#define QR_ACLK(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0) { long long now_ = clock64(); g_qr_aclk[slot] += now_ - aclk_; aclk_ = now_; } } while (0)
#define QR_CLK(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0) { long long now_ = clock64(); g_qr_clk[slot] += now_ - tclk_; tclk_ = now_; } } while (0)
#else
#define QR_CLK(slot) do { } while (0)
#define QR_ACLK(slot) do { } while (0)
#endif

template <int RPT, bool KEEP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) qr_panel_kernel(double* __restrict__ W, int ld, int m, int j0, int nbp,
                                                               double* __restrict__ tau_out, double* __restrict__ T_out,
                                                               const double* __restrict__ Tprev) {
#ifdef TN_QR_TIMING
  long long tclk_ = clock64();
#endif
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), c = (int)cluster.block_rank();
  __shared__ double red[2][WARPS][QNB];        // per-warp partial Gram rows
  __shared__ double rowj[2][QNB];               // row j of the panel (CTA 0)
  __shared__ double slots[2][QMAXC][QSLOT];     // cluster exchange: [parity][source CTA][g(32) | row j(32)]
  __shared__ double Z[QNB * QNB];               // Z[k*32 + j] = V_k^T v_j (k < j)
  __shared__ double Ts[QNB * QNB];              // compact-WY factor
  __shared__ double taus[QNB];
  __shared__ uint64_t xbar[2];                  // arrival of the C contributions of a column (double-buffered like the slots)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int DI = (QNB / WARPS) < RPT ? (QNB / WARPS) : RPT;  // slots that may hold rows of the diagonal block (panel rows < 32)
  const int rows_total = m - j0;
  const int row_lo = c * (RPT * WARPS);        // first panel row of this CTA

  double x[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int pr = row_lo + i * WARPS + warp;  // panel row
    x[i] = (pr < rows_total && lane < nbp) ? W[(size_t)(j0 + pr) * ld + j0 + lane] : 0.0;
  }
  for (int idx = tid; idx < QNB * QNB; idx += (WARPS * 32)) {
    Z[idx] = 0.0;
    Ts[idx] = 0.0;
  }
  if (tid < QNB) taus[tid] = 0.0;
  if (tid == 0) {
    mbar_init_q(&xbar[0], 1);
    mbar_init_q(&xbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (C > 1) cluster_barrier();  // every CTA of the cluster is resident (and its mbarriers live) before remote stores start

  // ---- fused look-ahead: the block reflector of the PREVIOUS panel (columns [j0-32, j0), rows [j0-32, m), factor Tprev) is
  // applied to this panel's columns while they sit in the registers,  X <- X - V (Tprev^T (V^T X)),  instead of by a separate
  // launch on the critical path between two panels (~20 us of launch + latency chain per panel).  The 32 rows above the panel
  // (they become rows of R) are carried by CTA 0 as four extra row slots per thread.
  extern __shared__ __align__(16) double fz[];
  if constexpr (KEEP) if (Tprev != nullptr) {
    constexpr int ROWS = RPT * WARPS;                  // rows of this CTA
    double* Vs = fz;                                   // [ROWS][32]   previous reflectors, rows of this CTA (all below their diagonal block)
    double* Vt = Vs + ROWS * QNB;                      // [32][32]     unit-lower-triangular diagonal block (CTA 0)
    double* Tp = Vt + QNB * QNB;                       // [32][32]
    double* Wtot = Tp + QNB * QNB;                     // [32][32]     V^T X summed over the cluster
    double* Yn = Wtot + QNB * QNB;                     // [32][32]     -Tprev^T Wtot
    double* part = Yn + QNB * QNB;                     // [WARPS][8][32] per-warp partial rows of V^T X (one round of 8 reflectors)
    double* Wex = part + WARPS * 8 * QNB;              // [C][32][32]  cluster exchange
    const int jp = j0 - QNB;                           // first column / row of the previous panel
    double xt[4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
      xt[t] = (c == 0 && lane < nbp) ? W[(size_t)(jp + t * WARPS + warp) * ld + j0 + lane] : 0.0;
    {
      double v[RPT];
#pragma unroll
      for (int u = 0; u < RPT; ++u) {                  // element idx = r*32 + q, one deep batch of independent loads
        const int idx = u * (WARPS * 32) + tid, r = idx >> 5, q = idx & 31;
        v[u] = (row_lo + r < rows_total) ? W[(size_t)(j0 + row_lo + r) * ld + jp + q] : 0.0;
      }
      double vt4[4], tp4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = u * (WARPS * 32) + tid, r = idx >> 5, q = idx & 31;
        vt4[u] = (c == 0 && r > q) ? W[(size_t)(jp + r) * ld + jp + q] : (r == q ? 1.0 : 0.0);
        tp4[u] = Tprev[idx];
      }
#pragma unroll
      for (int u = 0; u < RPT; ++u) Vs[u * (WARPS * 32) + tid] = v[u];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        Vt[u * (WARPS * 32) + tid] = vt4[u];
        Tp[u * (WARPS * 32) + tid] = tp4[u];
      }
    }
    __syncthreads();
    // V^T X over this thread's rows: acc[q] = sum_i V[row_i, q] x[i]  (column `lane` of X); the V row is a shared-memory broadcast
    double acc[QNB];
#pragma unroll
    for (int q = 0; q < QNB; ++q) acc[q] = 0.0;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const double2* vr = reinterpret_cast<const double2*>(Vs + (i * WARPS + warp) * QNB);
#pragma unroll
      for (int q2 = 0; q2 < QNB / 2; ++q2) {
        const double2 vv = vr[q2];
        acc[2 * q2] = fma(vv.x, x[i], acc[2 * q2]);
        acc[2 * q2 + 1] = fma(vv.y, x[i], acc[2 * q2 + 1]);
      }
    }
    if (c == 0) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double2* vr = reinterpret_cast<const double2*>(Vt + (t * WARPS + warp) * QNB);
#pragma unroll
        for (int q2 = 0; q2 < QNB / 2; ++q2) {
          const double2 vv = vr[q2];
          acc[2 * q2] = fma(vv.x, xt[t], acc[2 * q2]);
          acc[2 * q2 + 1] = fma(vv.y, xt[t], acc[2 * q2 + 1]);
        }
      }
    }
    // sum over the warps (fixed order), 8 reflectors per round, and hand this CTA's partial to slot c of every CTA
#pragma unroll
    for (int rd = 0; rd < QNB / 8; ++rd) {
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) part[(warp * 8 + qq) * QNB + lane] = acc[rd * 8 + qq];
      __syncthreads();
      {
        double sacc = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < WARPS; ++w2) sacc += part[w2 * 8 * QNB + tid];   // tid = qq*32 + k
        const int e = rd * 8 * QNB + tid;
        if (C == 1) {
          Wex[e] = sacc;
        } else {
          for (int dst = 0; dst < C; ++dst) cluster.map_shared_rank(Wex, dst)[c * QNB * QNB + e] = sacc;
        }
      }
      __syncthreads();
    }
    if (C > 1) cluster_barrier();
    for (int e = tid; e < QNB * QNB; e += WARPS * 32) {
      double sacc = 0.0;
      for (int src = 0; src < C; ++src) sacc += Wex[src * QNB * QNB + e];
      Wtot[e] = sacc;
    }
    __syncthreads();
    for (int e = tid; e < QNB * QNB; e += WARPS * 32) {
      const int i = e >> 5, cc = e & 31;
      double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < QNB; ++k) s4[k & 3] = fma(Tp[k * QNB + i], Wtot[k * QNB + cc], s4[k & 3]);
      Yn[e] = -((s4[0] + s4[1]) + (s4[2] + s4[3]));
    }
    __syncthreads();
    // X += V Yn
#pragma unroll
    for (int q = 0; q < QNB; ++q) acc[q] = Yn[q * QNB + lane];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const double2* vr = reinterpret_cast<const double2*>(Vs + (i * WARPS + warp) * QNB);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int q2 = 0; q2 < QNB / 2; ++q2) {
        const double2 vv = vr[q2];
        s0 = fma(vv.x, acc[2 * q2], s0);
        s1 = fma(vv.y, acc[2 * q2 + 1], s1);
      }
      x[i] += s0 + s1;
    }
    if (c == 0) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double2* vr = reinterpret_cast<const double2*>(Vt + (t * WARPS + warp) * QNB);
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < QNB / 2; ++q2) {
          const double2 vv = vr[q2];
          s0 = fma(vv.x, acc[2 * q2], s0);
          s1 = fma(vv.y, acc[2 * q2 + 1], s1);
        }
        if (lane < nbp) W[(size_t)(jp + t * WARPS + warp) * ld + j0 + lane] = xt[t] + (s0 + s1);
      }
    }
    // columns >= nbp of a narrow last panel carried zeros and stay zero (x = 0 there, Yn column 0)
  }
  QR_CLK(0);

  double myscale = 0.0;  // 1 / (alpha - beta) of column `lane`: the reflector tails stay unscaled in the registers until the write-back
  for (int j = 0; j < nbp; ++j) {
    const int par = j & 1;
    // ---- row j of the panel Gram matrix, g_k = sum_{rows >= j} P[r,j] P[r,k], over this thread's rows ----
    // only the rows of the diagonal block (panel rows < 32: slots i < DI of CTA 0) can lie above row j; all others take the plain FMA
    double xjs[KEEP ? RPT : 1];  // column j of this thread's rows (one warp shuffle per row; reused by the update when KEEP)
    {
      // FP64 results come back after ~38 cycles on this part: 8 independent accumulators and a tree keep the chain short
      double acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = 0.0;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const double xj = __shfl_sync(0xffffffffu, x[i], j);
        if (KEEP) xjs[i] = xj;
        double xi = x[i];
        if (i < DI) xi = (row_lo + i * WARPS + warp >= j) ? xi : 0.0;
        acc[i & 7] = fma(xj, xi, acc[i & 7]);
      }
      red[par][warp][lane] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
      if (c == 0 && warp == (j & (WARPS - 1))) {  // row j of the panel = slot i = j / WARPS of this warp
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < DI; ++i) v = (i * WARPS + warp == j) ? x[i] : v;
        rowj[par][lane] = v;
      }
    }
    QR_CLK(1);
    __syncthreads();
    QR_CLK(2);
    double gk = 0.0, gj = 0.0, rjk, alpha;
    if (C == 1) {
      double pk[WARPS], pj[WARPS];
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        pk[w] = red[par][w][lane];
        pj[w] = red[par][w][j];
      }
#pragma unroll
      for (int st = 1; st < WARPS; st *= 2)
#pragma unroll
        for (int w = 0; w + st < WARPS; w += 2 * st) {
          pk[w] += pk[w + st];
          pj[w] += pj[w + st];
        }
      gk = pk[0];
      gj = pj[0];
      rjk = rowj[par][lane];
      alpha = rowj[par][j];
    } else {
      if (warp == 0) {
        double pk[WARPS];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) pk[w] = red[par][w][lane];
#pragma unroll
        for (int st = 1; st < WARPS; st *= 2)
#pragma unroll
          for (int w = 0; w + st < WARPS; w += 2 * st) pk[w] += pk[w + st];
        const double g = pk[0];
        const double rj = (c == 0) ? rowj[par][lane] : 0.0;
        // this CTA's partial row goes to slot c of every CTA (itself included) with asynchronous remote stores that count on
        // the destination's mbarrier: no cluster-wide barrier and no fence on the per-column path (barrier.cluster with release
        // semantics costs a GPU-scope MEMBAR plus an L1 invalidate, ~1600 cycles per column)
        if (lane == 0) mbar_expect_tx_q(&xbar[par], (unsigned)(C * QSLOT * sizeof(double)));
        for (int dst = 0; dst < C; ++dst) {
          st_async_q(&slots[par][c][lane], g, &xbar[par], (unsigned)dst);
          st_async_q(&slots[par][c][QNB + lane], rj, &xbar[par], (unsigned)dst);
        }
      }
      mbar_wait_q(&xbar[par], (unsigned)((j >> 1) & 1));
      for (int s = 0; s < C; ++s) {  // fixed order: identical in every CTA and from run to run
        gk += slots[par][s][lane];
        gj += slots[par][s][j];
      }
      rjk = slots[par][0][QNB + lane];
      alpha = slots[par][0][QNB + j];
    }
    QR_CLK(3);
    double tau, beta, scale;
    // nothing below the diagonal, or a column that is zero to the underflow threshold (its sum of squares would be formed from
    // denormal products: exactly dependent columns shrink by 1e-16 per reflector under FMA and get there): H = 1, as dlarfg
    // does for a zero tail; the sub-diagonal entries are cleared (scale = 0), a perturbation of 1e-140 of the input
    if (rows_total - 1 - j == 0 || !(gj > 1e-280)) {
      tau = 0.0; beta = alpha; scale = 0.0;
    } else {
      // beta = -sign(alpha) |x|, tau = (beta - alpha) / beta = 1 + |alpha| / |x|, scale = 1 / (alpha - beta): one rsqrt and one
      // reciprocal instead of a square root and two divisions on the critical path of the column
      const double rn = rsqrt(gj);            // 1 / |x|
      const double nrm = gj * rn;
      const double aabs = fabs(alpha);
      beta = alpha >= 0.0 ? -nrm : nrm;
      tau = fma(aabs, rn, 1.0);
      const double inv = __drcp_rn(aabs + nrm);
      scale = alpha >= 0.0 ? inv : -inv;
    }
    // k > j: v^T a_k;  k < j: V_k^T v_j (column k holds the unscaled tail u_k = v_k / myscale_k)
    const double y = (gk - beta * rjk) * scale * (lane < j ? myscale : 1.0);
    const double ty = tau * y;
    if (lane == j) myscale = scale;
    // ---- apply H_j to the columns k > j (registers only), store v below the diagonal and beta on it ----
    // x[r,k] -= (tau y_k scale) x[r,j] for k > j: the coefficient is zero in the lanes k <= j (column j keeps its unscaled tail), so
    // rows below the diagonal block need exactly one FMA and no branch
    const double coef = lane > j ? ty * scale : 0.0;
    QR_CLK(4);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const double xj = KEEP ? xjs[i] : __shfl_sync(0xffffffffu, x[i], j);
      if (i < DI) {
        const int pr = row_lo + i * WARPS + warp;
        if (pr > j) {
          x[i] = fma(-coef, xj, x[i]);
        } else if (pr == j) {
          if (lane > j) x[i] -= ty;
          else if (lane == j) x[i] = beta;
        }
      } else {
        x[i] = fma(-coef, xj, x[i]);
      }
    }
    if (c == 0 && warp == (j & (WARPS - 1))) {  // the warp that owns row j records the T-factor inputs
      if (lane < j) Z[lane * QNB + j] = y;
      if (lane == j) taus[j] = tau;
    }
    QR_CLK(5);
  }
  __syncthreads();

  // ---- compact-WY factor.  T^-1 = S with S_jj = 1 / tau_j and S_ij = V_i^T V_j = Z[i][j] for i < j (from T^-1 + T^-T = V^T V),
  // so T is the inverse of a 32 x 32 upper-triangular matrix: recursive doubling over diagonal blocks of size 1, 2, 4, 8, 16,
  //   [[A, B], [0, D]]^-1 = [[A^-1, -A^-1 B D^-1], [0, D^-1]],
  // every level a batch of tiny products done by the whole CTA (the column-by-column dlarft recurrence is a chain of 32
  // dependent steps, 16 us at FP64 latencies).  A reflector with tau = 0 (H = 1) decouples: S_jj = 1, row / column j of Z zero,
  // T_jj = 0 afterwards.
  if (c == 0) {
    const int nt = WARPS * 32;
    for (int idx = tid; idx < QNB * QNB; idx += nt) {
      const int i = idx >> 5, j = idx & 31;
      const bool off = i < nbp && j < nbp && taus[i] != 0.0 && taus[j] != 0.0;
      Ts[idx] = (i == j) ? ((i < nbp && taus[i] != 0.0) ? taus[i] : 1.0) : ((i < j && off) ? Z[idx] : 0.0);  // diag: A^-1 of the 1x1 blocks
    }
    __syncthreads();
    for (int sz = 1; sz < QNB; sz *= 2) {
      // block pair p: A = rows/cols [2p*sz, 2p*sz+sz), D = [2p*sz+sz, 2p*sz+2sz), B = A-rows x D-cols (still holds S)
      // step 1: Y = B D^-1 (kept in Z as scratch);  step 2: X = -A^-1 Y, written over B
      for (int e = tid; e < (QNB / (2 * sz)) * sz * sz; e += nt) {
        const int p = e / (sz * sz), r = (e / sz) % sz, q = e % sz;
        const int a0 = 2 * p * sz, d0 = a0 + sz;
        double acc0 = 0.0, acc1 = 0.0;
        for (int l = 0; l <= q; ++l) {  // D^-1 is upper triangular
          const double pr = Ts[(a0 + r) * QNB + d0 + l] * Ts[(d0 + l) * QNB + d0 + q];
          if (l & 1) acc1 += pr; else acc0 += pr;
        }
        Z[(a0 + r) * QNB + d0 + q] = acc0 + acc1;
      }
      __syncthreads();
      for (int e = tid; e < (QNB / (2 * sz)) * sz * sz; e += nt) {
        const int p = e / (sz * sz), r = (e / sz) % sz, q = e % sz;
        const int a0 = 2 * p * sz, d0 = a0 + sz;
        double acc0 = 0.0, acc1 = 0.0;
        for (int l = r; l < sz; ++l) {  // A^-1 is upper triangular
          const double pr = Ts[(a0 + r) * QNB + a0 + l] * Z[(a0 + l) * QNB + d0 + q];
          if (l & 1) acc1 += pr; else acc0 += pr;
        }
        Ts[(a0 + r) * QNB + d0 + q] = -(acc0 + acc1);
      }
      __syncthreads();
    }
    for (int idx = tid; idx < QNB * QNB; idx += nt) {
      const int i = idx >> 5, j = idx & 31;
      const bool dead = (i == j) && !(i < nbp && taus[i] != 0.0);
      T_out[idx] = dead ? 0.0 : Ts[idx];
    }
    if (tid < nbp) tau_out[j0 + tid] = taus[tid];
  }
  QR_CLK(6);
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int pr = row_lo + i * WARPS + warp;
    if (pr < rows_total && lane < nbp) W[(size_t)(j0 + pr) * ld + j0 + lane] = pr > lane ? x[i] * myscale : x[i];  // v below the diagonal
  }
  if (C > 1) cluster_barrier();  // no CTA exits while a peer may still address its shared memory
  QR_CLK(7);
}

#ifdef TN_QR_TIMING
extern "C" int tn_qr_debug_clocks(long long* out8, int reset) {   // out8: 16 values, panel phases then apply phases
  long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (out8) {
    if (cudaMemcpyFromSymbol(h, g_qr_clk, sizeof(h)) != cudaSuccess) return -2;
    for (int i = 0; i < 8; ++i) out8[i] = h[i];
    if (cudaMemcpyFromSymbol(h, g_qr_aclk, sizeof(h)) != cudaSuccess) return -2;
    for (int i = 0; i < 8; ++i) out8[8 + i] = h[i];
  }
  if (reset) {
    long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(g_qr_clk, z, sizeof(z)) != cudaSuccess) return -2;
    if (cudaMemcpyToSymbol(g_qr_aclk, z, sizeof(z)) != cudaSuccess) return -2;
  }
  return 0;
}
static int qr_debug_skip() { const char* e = getenv("TNALG_QR_DEBUG_SKIP"); return e ? atoi(e) : 0; }
#else
static int qr_debug_skip() { return 0; }
#endif

// Cm[j0:m, c_begin:c_end] <- (1 - V op(T) V^T) Cm[j0:m, c_begin:c_end];  transT != 0: op(T) = T^T.
// One cluster per strip of QNC = 32 columns; the CTAs of the cluster split the rows (RB each).  The kernel is a chain of
// latencies (it moves ~100 KB and does ~1 MFLOP per CTA), so every global access is made once and in one deep batch: the
// unit-lower-trapezoidal reflector rows AND the CTA's block of C are staged in shared memory by 32 independent loads per
// thread, both GEMMs then run from shared memory and the result goes back with plain stores.
// Phase 1 (W = V^T C): warp w owns the 8 x 16 output tile (reflectors 8*(w%4).., columns 16*(w/4)..) over all rows of the CTA
// -- no cross-warp reduction; the CTAs' partial tiles meet through distributed shared memory and are summed in CTA order
// (deterministic).  Phase 2: C -= V (op(T) W).
// vp / cp: pitches of the staged reflector / C rows (36: conflict-free fragment reads, 32: when the padded blocks would not
// fit); cp = 0: C is read from global / L2 in both phases; vp = 0: neither block is staged (very tall panels).
__global__ void __launch_bounds__(ATHREADS, 1) qr_apply_kernel(const double* __restrict__ Wv, int ldv, int m, int j0, int nbp,
                                                               const double* __restrict__ T, int transT, double* __restrict__ Cm,
                                                               int ldc, int c_begin, int c_end, int RB, int vp, int cp) {
#ifdef TN_QR_TIMING
  long long aclk_ = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) g_qr_aclk[7] += 1;
#endif
  cg::cluster_group cluster = cg::this_cluster();
  const int CR = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
  extern __shared__ __align__(16) double sm_apply[];
  constexpr int WSZ = QNB * QNC;                            // 32 x 32 block of V^T C
  double* Tsm = sm_apply;                                   // 32 x 32
  double* Wtot = Tsm + QNB * QNB;                           // WSZ: sum over the CTAs
  double* W2 = Wtot + WSZ;                                  // WSZ: -op(T) V^T C
  double* Wex = W2 + WSZ;                                   // [CR][WSZ] cluster exchange
  double* Vs = Wex + (size_t)CR * WSZ;                      // RB x vp reflector rows of this CTA
  double* Cs = Vs + (size_t)RB * vp;                        // RB x cp block of C
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int strip = blockIdx.x / CR;
  const int c0 = c_begin + strip * QNC;
  const int row0 = j0 + cr * RB;                            // first global row of this CTA
  const int nrows = max(0, min(RB, m - row0));
#pragma unroll
  for (int u = 0; u < QNB * QNB / ATHREADS; ++u) Tsm[tid + u * ATHREADS] = T[tid + u * ATHREADS];
  auto vglobal = [&](int r, int k) -> double {  // unit-lower-trapezoidal reflector entry of local row r (not staged)
    const int rl = row0 + r - j0;
    if (r >= nrows || k >= nbp || rl < k) return 0.0;
    return rl == k ? 1.0 : Wv[(size_t)(row0 + r) * ldv + j0 + k];
  };
  auto cglobal = [&](int r, int cc) -> double { return (r < nrows && cc < c_end) ? Cm[(size_t)(row0 + r) * ldc + cc] : 0.0; };
  // staging: element idx = r*32 + k of both blocks; 16 + 16 independent loads per thread and pass
  for (int base = 0; vp > 0 && base < nrows * QNB; base += 16 * ATHREADS) {
    double v[16], cv[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = base + u * ATHREADS + tid;
      const int r = idx >> 5, k = idx & 31;
      const bool in = idx < nrows * QNB;
      v[u] = (in && k < nbp && (row0 + r - j0) > k) ? Wv[(size_t)(row0 + r) * ldv + j0 + k] : 0.0;
      cv[u] = (cp > 0 && in && c0 + k < c_end) ? Cm[(size_t)(row0 + r) * ldc + c0 + k] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = base + u * ATHREADS + tid;
      const int r = idx >> 5, k = idx & 31;
      if (idx < nrows * QNB) {
        Vs[r * vp + k] = (k < nbp && (row0 + r - j0) == k) ? 1.0 : v[u];
        if (cp > 0) Cs[r * cp + k] = cv[u];
      }
    }
  }
  QR_ACLK(0);
  if (CR > 1) cluster_barrier();  // peers are resident (remote writes below); doubles as the CTA barrier
  else __syncthreads();
  QR_ACLK(1);

  // ---- phase 1: tile (mt, 2 n-tiles) of W = V^T C over all rows of this CTA; 4 interleaved accumulators per tile ----
  {
    const int mt = warp & 3, nh = warp >> 2;                // reflectors 8*mt.., columns 16*nh..
    double acc[4][2][2];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) acc[u][nt][0] = acc[u][nt][1] = 0.0;
    const int n_quads = (nrows + 3) / 4;
    for (int q0 = 0; q0 < n_quads; q0 += 8) {
      double b[8][2], a[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = 4 * (q0 + u) + t;  // local row
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int kc = 16 * nh + 8 * nt + g;
          b[u][nt] = cp > 0 ? (r < nrows ? Cs[r * cp + kc] : 0.0) : cglobal(r, c0 + kc);
        }
        a[u] = vp > 0 ? (r < nrows ? Vs[r * vp + 8 * mt + g] : 0.0) : vglobal(r, 8 * mt + g);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) dmma884(acc[u & 3][nt], a[u], b[u][nt]);
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const double s0 = (acc[0][nt][0] + acc[1][nt][0]) + (acc[2][nt][0] + acc[3][nt][0]);
      const double s1 = (acc[0][nt][1] + acc[1][nt][1]) + (acc[2][nt][1] + acc[3][nt][1]);
      const int e = (8 * mt + g) * QNC + 16 * nh + 8 * nt + 2 * t;
      if (CR == 1) {
        Wex[e] = s0;
        Wex[e + 1] = s1;
      } else {
        for (int dst = 0; dst < CR; ++dst) {
          double* remote = cluster.map_shared_rank(Wex, dst) + cr * WSZ;
          remote[e] = s0;
          remote[e + 1] = s1;
        }
      }
    }
  }
  QR_ACLK(2);
  if (CR > 1) cluster_barrier(); else __syncthreads();
  QR_ACLK(3);
  for (int e = tid; e < WSZ; e += ATHREADS) {  // total over the CTAs in rank order
    double s = 0.0;
    for (int src = 0; src < CR; ++src) s += Wex[src * WSZ + e];
    Wtot[e] = s;
  }
  __syncthreads();
  for (int e = tid; e < WSZ; e += ATHREADS) {
    const int i = e / QNC, cc = e % QNC;
    // T is upper triangular and zero-padded: the full 32-term product with 4 independent accumulators (short FP64 chains)
    double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < QNB; ++k) s4[k & 3] = fma(transT ? Tsm[k * QNB + i] : Tsm[i * QNB + k], Wtot[k * QNC + cc], s4[k & 3]);
    W2[e] = -((s4[0] + s4[1]) + (s4[2] + s4[3]));
  }
  __syncthreads();
  QR_ACLK(4);
  // ---- phase 2: C += V W2 on this CTA's rows (M = rows, N = 32, K = 32; warps take groups of 8 rows) ----
  const int n_oct = (nrows + 7) / 8;
  constexpr int UO = 2;  // row groups per batch
  for (int o0 = warp; o0 < n_oct; o0 += UO * AWARPS) {
    double cacc[UO][QNC / 8][2];
#pragma unroll
    for (int u = 0; u < UO; ++u) {
      const int r = 8 * (o0 + u * AWARPS) + g;
#pragma unroll
      for (int nt = 0; nt < QNC / 8; ++nt) {
        const int kc = 8 * nt + 2 * t;
        if (cp > 0) {
          cacc[u][nt][0] = r < nrows ? Cs[r * cp + kc] : 0.0;
          cacc[u][nt][1] = r < nrows ? Cs[r * cp + kc + 1] : 0.0;
        } else {
          cacc[u][nt][0] = cglobal(r, c0 + kc);
          cacc[u][nt][1] = cglobal(r, c0 + kc + 1);
        }
      }
    }
#pragma unroll
    for (int ks = 0; ks < QNB / 4; ++ks) {
      double bw[QNC / 8];
#pragma unroll
      for (int nt = 0; nt < QNC / 8; ++nt) bw[nt] = W2[(4 * ks + t) * QNC + 8 * nt + g];
#pragma unroll
      for (int u = 0; u < UO; ++u) {
        const int r = 8 * (o0 + u * AWARPS) + g;
        const double a = vp > 0 ? (r < nrows ? Vs[r * vp + 4 * ks + t] : 0.0) : vglobal(r, 4 * ks + t);
#pragma unroll
        for (int nt = 0; nt < QNC / 8; ++nt) dmma884(cacc[u][nt], a, bw[nt]);
      }
    }
#pragma unroll
    for (int u = 0; u < UO; ++u) {
      const int r = 8 * (o0 + u * AWARPS) + g;
#pragma unroll
      for (int nt = 0; nt < QNC / 8; ++nt) {
        const int cc = c0 + 8 * nt + 2 * t;
        if (r < nrows && cc < c_end) Cm[(size_t)(row0 + r) * ldc + cc] = cacc[u][nt][0];
        if (r < nrows && cc + 1 < c_end) Cm[(size_t)(row0 + r) * ldc + cc + 1] = cacc[u][nt][1];
      }
    }
  }
  QR_ACLK(5);
  if (CR > 1) cluster_barrier();  // no CTA exits while a peer may still write its exchange slots
  QR_ACLK(6);
}

// out (rows x cols, ld ldo) = in^T or in;  tiled through shared memory
__global__ void qr_copy_kernel(const double* __restrict__ in, int ldi, double* __restrict__ out, int ldo, int rows_out, int cols_out,
                               int transpose, const double* __restrict__ scale) {
  __shared__ double tile[32][33];
  const double sc = scale ? *scale : 1.0;  // power of two: exact
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;  // bx: output column block, by: output row block
  if (!transpose) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int r = by + i, c = bx + threadIdx.x;
      if (r < rows_out && c < cols_out) out[(size_t)r * ldo + c] = sc * in[(size_t)r * ldi + c];
    }
    return;
  }
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // read in[c_out_block + i][r_out_block + x]
    const int rin = bx + i, cin = by + threadIdx.x;
    tile[i][threadIdx.x] = (rin < cols_out && cin < rows_out) ? in[(size_t)rin * ldi + cin] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = by + i, c = bx + threadIdx.x;
    if (r < rows_out && c < cols_out) out[(size_t)r * ldo + c] = sc * tile[threadIdx.x][i];
  }
}

// R (k x n) = upper trapezoid of the factored work matrix; Q (m x k) = [1; 0]
__global__ void qr_extract_r_kernel(const double* __restrict__ W, int ld, double* __restrict__ Rm, int k, int n, const double* __restrict__ scale) {
  const double sc = *scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)k * n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / n), c = (int)(i % n);
    Rm[i] = c >= r ? sc * W[(size_t)r * ld + c] : 0.0;
  }
}
__global__ void qr_eye_kernel(double* __restrict__ Q, int m, int k) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)m * k; i += (long long)gridDim.x * blockDim.x)
    Q[i] = (i / k == i % k) ? 1.0 : 0.0;
}

// in place: columns [0, k) of the factored work matrix become the explicit unit-lower-trapezoidal V (zeros above the diagonal, ones on
// it; a reflector with tau = 0 is the identity and its column is cleared).  R has been extracted before.
__global__ void qr_make_v_kernel(double* __restrict__ W, int ld, int k, const double* __restrict__ tau) {
  // only the upper triangle + diagonal of the leading k x k block changes
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)k * k; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / k), c = (int)(i % k);
    if (c > r) W[(size_t)r * ld + c] = 0.0;
    else if (c == r) W[(size_t)r * ld + c] = tau[c] != 0.0 ? 1.0 : 0.0;
  }
}
// S = T^-1 = diag(1 / tau) + striu(V^T V), from the partial Gram matrices Gp[s] (K split of the GEMM, summed in order);
// Tf = 0 except for the diagonal 32 x 32 blocks, which are the panel factors.  Both (kp x kp), identity / zero beyond k.
__global__ void qr_build_s_kernel(const double* __restrict__ Gp, int n_split, size_t split_stride, double* __restrict__ S,
                                  double* __restrict__ Tf, const double* __restrict__ Tall, int kp, int k) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)kp * kp; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / kp), c = (int)(i % kp);
    double g = 0.0;
    if (c > r && c < k && (r >> 5) != (c >> 5))
      for (int sidx = 0; sidx < n_split; ++sidx) g += Gp[sidx * split_stride + i];
    S[i] = g;   // only the blocks above the block diagonal are read
    Tf[i] = ((r >> 5) == (c >> 5) && c < k && r < k) ? Tall[(size_t)(r >> 5) * QNB * QNB + (r & 31) * QNB + (c & 31)] : 0.0;
  }
}

struct QrSide {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaEvent_t join2[2] = {nullptr, nullptr};   // fused look-ahead: wide updates of even / odd panels
};
static QrSide* qr_side() {
  static QrSide x;
  if (!x.s) {
    if (cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.join2[0], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.join2[1], cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &x;
}

static int grid_for(long long n) { return (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8); }

// dynamic shared memory of the fused look-ahead (qr_panel_kernel with Tprev): reflector rows, diagonal block, T, V^T X, its
// image under -T^T, one reduction round and the cluster exchange
static size_t panel_fused_smem(int rpt, int C) {
  return sizeof(double) * ((size_t)rpt * 8 * QNB + 4 * (size_t)QNB * QNB + 8 * 8 * QNB + (size_t)C * QNB * QNB);
}

template <int RPT, bool KEEP, int WARPS>
static int launch_panel_t(double* W, int ld, int m, int j0, int nbp, double* tau, double* T, const double* Tprev, int C,
                          cudaStream_t stream) {
  size_t smem = 0;
  if (Tprev) {
    TN_REQUIRE(KEEP && WARPS == 8 && j0 >= QNB, "tn_qr: internal: fused look-ahead on an unsupported panel configuration");
    smem = panel_fused_smem(RPT, C);
    static size_t configured = 0;
    if (smem > configured) {
      TN_CUDA(cudaFuncSetAttribute(qr_panel_kernel<RPT, KEEP, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(190 * 1024)));
      configured = 190 * 1024;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TN_CUDA(cudaLaunchKernelEx(&cfg, qr_panel_kernel<RPT, KEEP, WARPS>, W, ld, m, j0, nbp, tau, T, Tprev));
  TN_LAUNCHED();
  return TN_OK;
}

// Tprev != nullptr: the kernel first applies the previous panel's block reflector to its own columns (8-warp variants only: see
// panel_can_fuse)
static int launch_panel(double* W, int ld, int m, int j0, int nbp, double* tau, double* T, const double* Tprev, cudaStream_t stream) {
  const int rows = m - j0;
  // Up to 2048 rows: 8 warps per CTA, 256 rows per CTA (32 per thread, the column-j broadcasts of phase A kept for the update),
  // a cluster of 1 / 2 / 4 / 8 CTAs.  The per-column cost is instruction issue: few warps keep the redundant scalar work small.
  // Taller panels: 16 warps, up to 768 rows per CTA (6144 per cluster).
  // rows per CTA: the per-column cost is instruction issue inside the CTA (shuffles + FMAs over the thread's rows) plus one
  // cluster exchange whose latency does not depend on the cluster size, so small slices win: 128 rows per CTA measured
  // 10 % faster than 256 (profiles/r02_qr.md; a non-portable 16-CTA cluster for 2048 rows measured no further gain)
  static const int cta_rows = getenv("TNALG_QR_PANEL_CTA_ROWS") ? std::max(32, atoi(getenv("TNALG_QR_PANEL_CTA_ROWS"))) : 128;
  int C = 1;
  while (C < QMAXC && (rows + C - 1) / C > cta_rows && rows / (2 * C) >= QNB) C *= 2;
  const int per_cta = (rows + C - 1) / C;
  TN_REQUIRE(C == 1 || per_cta >= QNB, "tn_qr: internal panel split");
  if (per_cta <= 256) {
    const int rpt = (per_cta + 7) / 8;
    if (rpt <= 4) return launch_panel_t<4, true, 8>(W, ld, m, j0, nbp, tau, T, Tprev, C, stream);
    if (rpt <= 8) return launch_panel_t<8, true, 8>(W, ld, m, j0, nbp, tau, T, Tprev, C, stream);
    if (rpt <= 16) return launch_panel_t<16, true, 8>(W, ld, m, j0, nbp, tau, T, Tprev, C, stream);
    return launch_panel_t<32, true, 8>(W, ld, m, j0, nbp, tau, T, Tprev, C, stream);
  }
  const int rpt = (per_cta + 15) / 16;
  TN_REQUIRE(rpt <= QMAX_RPT, "tn_qr: %d rows exceed the panel capacity (%d rows)", rows, QMAXC * QMAX_RPT * 16);
  if (rpt <= 32) return launch_panel_t<32, false, 16>(W, ld, m, j0, nbp, tau, T, Tprev, C, stream);
  return launch_panel_t<QMAX_RPT, false, 16>(W, ld, m, j0, nbp, tau, T, Tprev, C, stream);
}

// the panel of `rows` rows runs on an 8-warp variant (<= 256 rows per CTA of a cluster of <= 8): the fused look-ahead is available
static bool panel_can_fuse(int rows) {
  if (rows < QNB) return false;
  return (rows + QMAXC - 1) / QMAXC <= 256;
}

static int launch_apply(const double* Wv, int ldv, int m, int j0, int nbp, const double* T, int transT, double* Cm, int ldc,
                        int c_begin, int c_end, cudaStream_t stream) {
  if (c_end <= c_begin || m - j0 <= 0) return TN_OK;
  const int rows = m - j0;
  const int strips = (c_end - c_begin + QNC - 1) / QNC;
  static const int apply_rows = getenv("TNALG_QR_APPLY_CTA_ROWS") ? std::max(32, atoi(getenv("TNALG_QR_APPLY_CTA_ROWS"))) : 128;   // 128-row slices: one staging pass per CTA (-5 % at 512 x 256)
  int CR = 1;
  while (CR < QMAXC && (rows + CR - 1) / CR > apply_rows) CR *= 2;
  const int RB = ((rows + CR - 1) / CR + 7) / 8 * 8;
  const size_t fixed = sizeof(double) * ((size_t)QNB * QNB + (size_t)CR * QNB * QNC + 2 * (size_t)QNB * QNC);
  const size_t cap = 227 * 1024;
  auto fits = [&](int v, int c) { return fixed + sizeof(double) * (size_t)RB * (size_t)(v + c) <= cap; };
  int vp = AVP, cp = AVP;                                 // both blocks staged with padded rows
  if (!fits(vp, cp)) { vp = QNB; cp = QNB; }              // ... unpadded
  if (!fits(vp, cp)) { vp = AVP; cp = 0; }                // reflectors only
  if (!fits(vp, cp)) vp = QNB;
  if (!fits(vp, cp)) vp = 0;                              // too tall to stage: both operands are read from global / L2
  const size_t smem = fixed + sizeof(double) * (size_t)RB * (size_t)(vp + cp);
  static size_t configured = 0;
  if (smem > configured) {
    TN_CUDA(cudaFuncSetAttribute(qr_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(strips * CR);
  cfg.blockDim = dim3(ATHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CR;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TN_CUDA(cudaLaunchKernelEx(&cfg, qr_apply_kernel, Wv, ldv, m, j0, nbp, T, transT, Cm, ldc, c_begin, c_end, RB, vp, cp));
  TN_LAUNCHED();
  return TN_OK;
}

// ---- Q from the compact-WY form of ALL reflectors:  Q = [1; 0] - V (T V1^T),  T^-1 = diag(1 / tau) + striu(V^T V) ----
// Applying the panels one after the other to the identity is a chain of min(m,n)/32 latency-bound launches (1.4 ms of the
// 4.4 ms of a 2048 x 1024 factorisation); here the same product is three large DMMA chain GEMMs plus the inversion of the
// upper-triangular T^-1 by recursive doubling over its diagonal blocks (the 32 x 32 ones are the panel factors),
//   [[A, B], [0, D]]^-1 = [[A^-1, -A^-1 B D^-1], [0, D^-1]],   two batched GEMMs per level.
// All GEMMs run in the deterministic schedule (whole tiles, fixed summation order); the K range of the Gram product is cut
// into separate outputs that are added in order, so the result is bit-reproducible like the rest of the factorisation.
constexpr int kQrGemmSlots = 16;   // descriptor chunks in the workspace: 1 Gram + 2 per level (<= 6 levels) + 2
static int qr_kp(int k) {
  const int panels = (k + QNB - 1) / QNB;
  int p2 = 1;
  while (p2 < panels) p2 *= 2;
  return p2 * QNB;
}
static int qr_gram_split(int m, int k) {
  const int tiles = ((k + kTileBM - 1) / kTileBM) * ((k + kTileBN - 1) / kTileBN);
  int ns = 1;
  while (ns < 4 && tiles * ns * 2 <= 2 * sm_count() && m % (2 * ns * kBK) == 0 && m / (2 * ns) >= 256) ns *= 2;
  return ns;
}
static int qr_fullt_min_panels() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TNALG_QR_WY_MIN_PANELS");
    v = e ? atoi(e) : 8;
  }
  return v;
}
static size_t qr_gemm_slot_bytes() { return align_up(tn_chain_gemm_workspace_bytes(32, 32)); }

// The GEMMs of the Q formation are tiny and strictly ordered: their descriptors (problem / link tables of the chain GEMM) hold
// only workspace addresses, so ALL of them are built on the host first and reach the device in ONE copy; the launches then
// follow each other without a host-to-device transfer in between (ten calls of tn_chain_gemm cost two copies each, ~80 us of a
// 0.6 ms factorisation at 512 x 256).
struct QrGemmBatch {
  struct Call {
    GemmLaunch L;
    GemmSchedule S;
  };
  std::vector<Call> calls;
  std::vector<char> host;
  char* dev = nullptr;
  size_t slot_bytes = 0, link_off = 0;
  std::vector<ProblemDev> ph;
  std::vector<LinkDev> lh;

  explicit QrGemmBatch(char* dev_slots) : dev(dev_slots) {
    slot_bytes = qr_gemm_slot_bytes();
    link_off = align_up(sizeof(ProblemDev) * 32);
  }
  void begin() { ph.clear(); lh.clear(); }
  void add(const double* A, const double* B, double* Cm, double alpha, int accumulate) {
    LinkDev l = {};
    l.A = A; l.B = B; l.has_op = 0;
    ProblemDev q = {};
    q.C = Cm; q.alpha = alpha; q.link_begin = (int)lh.size(); q.link_count = 1; q.accumulate = accumulate;
    lh.push_back(l);
    ph.push_back(q);
  }
  // closes the call opened by begin(); returns its index (-1: no block in it), or a negative status - 100 on error
  int end(int mode, int M, int N, int K, int lda, int ldb, int ldc) {
    if (ph.empty()) return -1;
    if (ph.size() > 32 || (int)calls.size() >= kQrGemmSlots) { set_error("tn_qr: too many blocks / products in the Q formation"); return -100 + TN_ERR_INVALID; }
    Call c;
    c.L = GemmLaunch{mode, M, N, K, 1, lda, ldb, ldc, (int)ph.size(), (int)lh.size(), 1, 1};
    const int st = gemm_plan_schedule(c.L, ph.data(), lh.data(), nullptr, nullptr, &c.S);
    if (st != TN_OK) return -100 + st;
    const size_t base = calls.size() * slot_bytes;
    host.resize(base + slot_bytes, 0);
    memcpy(host.data() + base, ph.data(), sizeof(ProblemDev) * ph.size());
    memcpy(host.data() + base + link_off, lh.data(), sizeof(LinkDev) * lh.size());
    calls.push_back(c);
    return (int)calls.size() - 1;
  }
  int upload(cudaStream_t stream) {
    if (host.empty()) return TN_OK;
    TN_CUDA(cudaMemcpyAsync(dev, host.data(), host.size(), cudaMemcpyHostToDevice, stream));
    return TN_OK;
  }
  int launch(int i, cudaStream_t stream) {
    if (i < 0) return TN_OK;
    const Call& c = calls[(size_t)i];
    return gemm_launch(c.L, c.S, reinterpret_cast<const ProblemDev*>(dev + (size_t)i * slot_bytes),
                       reinterpret_cast<const LinkDev*>(dev + (size_t)i * slot_bytes + link_off), nullptr, nullptr, 1.0, stream);
  }
};

// W: factored work matrix (V below the diagonal of its first k columns, R already extracted), Qdst (m x k) receives Q
static int form_q_compact_wy(double* W, int ld, int m, int k, const double* tau, const double* Tall, double* Qdst, double* Gp,
                             double* S, double* Tf, double* Y, char* slots, cudaStream_t stream) {
  const int kp = qr_kp(k);
  const int ns = qr_gram_split(m, k);
  const size_t kk = (size_t)kp * kp;
  QrGemmBatch gb(slots);
#define TN_QR_END(var, ...)                          \
  const int var = gb.end(__VA_ARGS__);               \
  if (var <= -100) return var + 100;
  gb.begin();   // partial Gram matrices over row ranges of V
  const int rows_per = m / ns;
  for (int i = 0; i < ns; ++i) gb.add(W + (size_t)i * rows_per * ld, W + (size_t)i * rows_per * ld, Gp + i * kk, 1.0, 0);
  TN_QR_END(c_gram, TN_TN, k, k, rows_per, ld, ld, kp);
  std::vector<int> c_levels;
  for (int sz = QNB; sz < kp; sz *= 2) {
    gb.begin();
    for (int a0 = 0; a0 + sz < k; a0 += 2 * sz) {
      const int d0 = a0 + sz;
      gb.add(S + (size_t)a0 * kp + d0, Tf + (size_t)d0 * kp + d0, Y + (size_t)a0 * kp + d0, 1.0, 0);       // Y = B D^-1
    }
    TN_QR_END(cy, TN_NN, sz, sz, sz, kp, kp, kp);
    gb.begin();
    for (int a0 = 0; a0 + sz < k; a0 += 2 * sz) {
      const int d0 = a0 + sz;
      gb.add(Tf + (size_t)a0 * kp + a0, Y + (size_t)a0 * kp + d0, Tf + (size_t)a0 * kp + d0, -1.0, 0);     // -A^-1 Y
    }
    TN_QR_END(cx, TN_NN, sz, sz, sz, kp, kp, kp);
    c_levels.push_back(cy);
    c_levels.push_back(cx);
  }
  gb.begin();   // Y = T V1^T (k x k), V1 = leading k x k block of V
  gb.add(Tf, W, Y, 1.0, 0);
  TN_QR_END(c_tv, TN_NT, k, k, k, kp, ld, kp);
  gb.begin();
  gb.add(W, Y, Qdst, -1.0, 1);
  TN_QR_END(c_q, TN_NN, m, k, k, ld, kp, k);
#undef TN_QR_END
  TN_CHECK(gb.upload(stream));
  qr_make_v_kernel<<<grid_for((long long)k * k), 256, 0, stream>>>(W, ld, k, tau);
  TN_LAUNCHED();
  TN_CHECK(gb.launch(c_gram, stream));
  qr_build_s_kernel<<<grid_for((long long)kk), 256, 0, stream>>>(Gp, ns, kk, S, Tf, Tall, kp, k);
  TN_LAUNCHED();
  for (int c : c_levels) TN_CHECK(gb.launch(c, stream));
  TN_CHECK(gb.launch(c_tv, stream));
  qr_eye_kernel<<<grid_for((long long)m * k), 256, 0, stream>>>(Qdst, m, k);
  TN_LAUNCHED();
  TN_CHECK(gb.launch(c_q, stream));
  return TN_OK;
}

}  // namespace tn

using namespace tn;

extern "C" size_t tn_qr_workspace_bytes(int m, int n) {
  const int k = std::min(m, n);
  const size_t panels = (size_t)(k + QNB - 1) / QNB;
  size_t wy = 0;   // compact-WY formation of Q: partial Gram matrices, S, T, scratch (kp x kp each) + GEMM descriptor chunks
  if ((int)panels >= qr_fullt_min_panels()) {
    const size_t kk = (size_t)qr_kp(k) * qr_kp(k);
    wy = align_up(sizeof(double) * kk * qr_gram_split(m, k)) + 3 * align_up(sizeof(double) * kk) + kQrGemmSlots * qr_gemm_slot_bytes();
  }
  return align_up(sizeof(double) * (size_t)m * n) + align_up(sizeof(double) * (size_t)m * k) + align_up(sizeof(double) * panels * QNB * QNB) +
         align_up(sizeof(double) * (size_t)(k + QNB)) + 2 * 256 + 1024 + wy;
}

// trans_in != 0: the input is A^T, stored (n, m) row-major (the right-to-left move factorises the transposed matricisation
// without materialising it in the caller);  trans_q != 0: Q is written transposed, (k, m) row-major.
extern "C" int tn_qr_householder(const double* A, int m, int n, int trans_in, double* Q, int trans_q, double* R, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(A && Q && R && m > 0 && n > 0, "tn_qr_householder: bad arguments");
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_qr_householder: workspace must be 256-byte aligned");
  if (workspace_bytes < tn_qr_workspace_bytes(m, n)) {
    set_error("tn_qr_householder: workspace %zu < %zu bytes", workspace_bytes, tn_qr_workspace_bytes(m, n));
    return TN_ERR_WORKSPACE;
  }
  const int k = std::min(m, n);
  const int panels = (k + QNB - 1) / QNB;
  Carver cw(workspace, workspace_bytes);
  double* W = cw.take<double>((size_t)m * n);
  double* Qw = cw.take<double>((size_t)m * k);
  double* Tall = cw.take<double>((size_t)panels * QNB * QNB);
  double* tau = cw.take<double>((size_t)(k + QNB));
  double* scale2 = cw.take<double>(2);
  unsigned long long* slot = cw.take<unsigned long long>(1);
  TN_REQUIRE(W && Qw && Tall && tau && scale2 && slot, "tn_qr_householder: workspace carve failed");
  const bool wy = panels >= qr_fullt_min_panels();
  double *Gp = nullptr, *Sm = nullptr, *Tf = nullptr, *Ym = nullptr;
  char* gslots = nullptr;
  if (wy) {
    const size_t kk = (size_t)qr_kp(k) * qr_kp(k);
    Gp = cw.take<double>(kk * qr_gram_split(m, k));
    Sm = cw.take<double>(kk);
    Tf = cw.take<double>(kk);
    Ym = cw.take<double>(kk);
    gslots = cw.take<char>(kQrGemmSlots * qr_gemm_slot_bytes());
    TN_REQUIRE(Gp && Sm && Tf && Ym && gslots, "tn_qr_householder: workspace carve failed");
  }
  const dim3 tb(32, 8);
  // The column norms are formed as plain sums of squares, so the work copy is A scaled by a power of two to max|a_ij| in [1/2, 1)
  // (exact; R is scaled back): un-normalised MPS tensors reach 1e300 on long chains and their squares would overflow
  TN_CUDA(cudaMemsetAsync(slot, 0, sizeof(unsigned long long), stream));
  TN_CHECK(launch_pow2_scale(A, (long long)m * n, scale2, slot, stream));
  qr_copy_kernel<<<dim3((n + 31) / 32, (m + 31) / 32), tb, 0, stream>>>(A, trans_in ? m : n, W, n, m, n, trans_in ? 1 : 0, scale2);
  TN_LAUNCHED();
  // right-looking with look-ahead: after panel p, its reflectors are applied to the columns of panel p+1 first; the update of the
  // remaining columns then runs on a side stream while panel p+1 (latency bound, a handful of SMs) is factored on the main stream
  const int dbg_skip = qr_debug_skip();   // timing builds only: 1 = no Q formation, 2 = no wide trailing updates, 4 = no look-ahead updates, 8 = no panels
  QrSide* side = (panels > 2 && n > 4 * QNB && !getenv("TNALG_QR_NO_LOOKAHEAD")) ? qr_side() : nullptr;
  static const bool no_fuse = getenv("TNALG_QR_NO_FUSE") != nullptr;
  const bool fused = !no_fuse && !dbg_skip && panels >= 2 && panel_can_fuse(m - QNB);
  if (fused) {
    // Fused look-ahead: panel p applies the reflectors of panel p-1 to its OWN columns inside the panel kernel; the wide update
    // with the reflectors of panel p (columns from panel p+2 on; everything to the right for the last panel) runs on the side
    // stream concurrently with panel p+1.  Panel p needs the wide update p-2 (the last one that touched its columns).
    bool rec[2] = {false, false};
    for (int p = 0; p < panels; ++p) {
      const int j0 = p * QNB, nbp = std::min(QNB, k - j0);
      double* Tp = Tall + (size_t)p * QNB * QNB;
      if (side && p >= 2 && rec[p & 1]) TN_CUDA(cudaStreamWaitEvent(stream, side->join2[p & 1], 0));
      TN_CHECK(launch_panel(W, n, m, j0, nbp, tau, Tp, p > 0 ? Tp - QNB * QNB : nullptr, stream));
      const int nb_next = std::min(QNB, k - (j0 + nbp));   // the next panel takes exactly its own columns
      const int c_begin = (p + 1 < panels) ? j0 + nbp + nb_next : j0 + nbp;
      if (c_begin >= n) continue;
      if (side) {
        TN_CUDA(cudaEventRecord(side->fork, stream));
        TN_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
        TN_CHECK(launch_apply(W, n, m, j0, nbp, Tp, 1, W, n, c_begin, n, side->s));
        TN_CUDA(cudaEventRecord(side->join2[p & 1], side->s));
        rec[p & 1] = true;
      } else {
        TN_CHECK(launch_apply(W, n, m, j0, nbp, Tp, 1, W, n, c_begin, n, stream));
      }
    }
    if (side) {   // all wide updates done before R and Q are read (the side stream is in order: the last record covers them all)
      const int last = rec[(panels - 1) & 1] ? ((panels - 1) & 1) : ((panels - 2) & 1);
      if (rec[last]) TN_CUDA(cudaStreamWaitEvent(stream, side->join2[last], 0));
      if (rec[last ^ 1]) TN_CUDA(cudaStreamWaitEvent(stream, side->join2[last ^ 1], 0));
    }
  } else
  for (int p = 0; p < panels; ++p) {
    const int j0 = p * QNB, nbp = std::min(QNB, k - j0);
    const double* Tp = Tall + (size_t)p * QNB * QNB;
    if (!(dbg_skip & 8)) TN_CHECK(launch_panel(W, n, m, j0, nbp, tau, Tall + (size_t)p * QNB * QNB, nullptr, stream));
    const int next_end = std::min(n, j0 + nbp + QNB);
    if (!side || next_end >= n) {
      if (side && p > 0) TN_CUDA(cudaStreamWaitEvent(stream, side->join, 0));
      TN_CHECK(launch_apply(W, n, m, j0, nbp, Tp, 1, W, n, j0 + nbp, n, stream));
      if (side) side = nullptr;   // tail: plain ordering from here on
      continue;
    }
    if (p > 0) TN_CUDA(cudaStreamWaitEvent(stream, side->join, 0));   // the previous wide update has reached these columns
    if (!(dbg_skip & 4)) TN_CHECK(launch_apply(W, n, m, j0, nbp, Tp, 1, W, n, j0 + nbp, next_end, stream));
    TN_CUDA(cudaEventRecord(side->fork, stream));
    TN_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
    if (!(dbg_skip & 2)) TN_CHECK(launch_apply(W, n, m, j0, nbp, Tp, 1, W, n, next_end, n, side->s));
    TN_CUDA(cudaEventRecord(side->join, side->s));
    if (p == panels - 1) TN_CUDA(cudaStreamWaitEvent(stream, side->join, 0));   // wide matrix: the last update still runs on the side stream
  }
  qr_extract_r_kernel<<<grid_for((long long)k * n), 256, 0, stream>>>(W, n, R, k, n, scale2 + 1);
  TN_LAUNCHED();
  double* Qdst = trans_q ? Qw : Q;
  if (wy) {
    if (!(dbg_skip & 1)) TN_CHECK(form_q_compact_wy(W, n, m, k, tau, Tall, Qdst, Gp, Sm, Tf, Ym, gslots, stream));
  } else {
    qr_eye_kernel<<<grid_for((long long)m * k), 256, 0, stream>>>(Qdst, m, k);
    TN_LAUNCHED();
    for (int p = panels - 1; p >= 0 && !(dbg_skip & 1); --p) {
      const int j0 = p * QNB, nbp = std::min(QNB, k - j0);
      TN_CHECK(launch_apply(W, n, m, j0, nbp, Tall + (size_t)p * QNB * QNB, 0, Qdst, k, j0, k, stream));
    }
  }
  if (trans_q) {
    qr_copy_kernel<<<dim3((m + 31) / 32, (k + 31) / 32), tb, 0, stream>>>(Qw, k, Q, m, k, m, 1, nullptr);
    TN_LAUNCHED();
  }
  return TN_OK;
}

// Gauge moves on an MPS tensor T (a, d, b)  (SURVEY.md 8b):
//   tn_qr_l2r: T.reshape(a*d, b) = Q R        -> Q_out (a, d, k), R_out (k, b),  k = min(a*d, b)
//   tn_qr_r2l: T.reshape(a, d*b)^T = Q R      -> Q_out (k, d, b) = Q^T, R_out (k, a),  k = min(a, d*b)
extern "C" int tn_qr_l2r(const double* T, int a, int d, int b, double* Q_out, double* R_out, void* workspace, size_t workspace_bytes,
                         void* stream) {
  TN_REQUIRE(a > 0 && d > 0 && b > 0, "tn_qr_l2r: bad shape");
  return tn_qr_householder(T, a * d, b, 0, Q_out, 0, R_out, workspace, workspace_bytes, stream);
}
extern "C" int tn_qr_r2l(const double* T, int a, int d, int b, double* Q_out, double* R_out, void* workspace, size_t workspace_bytes,
                         void* stream) {
  TN_REQUIRE(a > 0 && d > 0 && b > 0, "tn_qr_r2l: bad shape");
  return tn_qr_householder(T, d * b, a, 1, Q_out, 1, R_out, workspace, workspace_bytes, stream);
}
