// Device-resident thick-restart Lanczos for the dominant eigenpair of (1 - tau*H_eff)  (a8, MPSClass.py:801-805).
//
// The Krylov space of (1 - tau*H) equals that of H, so the recurrence runs on H itself (tn_effh_matvec with
// (c_id, c_h) = (0, 1)): no 1e-4-shift cancellation, same Ritz vectors.  Per step:
//   w = H v_j                      chain-GEMM launches (FP64 tensor pipe) [+ caller's all-reduce when sharded]
//   h = V^T w ; w -= V h           twice (CGS2 full re-orthogonalisation), HBM-bound multi-dot / multi-axpy
//   beta = |w| ; v_{j+1} = w/beta  warp-shuffle dot, scale by a device scalar
// alpha/beta, the projected tridiagonal eigenproblem, the Ritz selection and the convergence flag live in a small
// device struct; the host only enqueues kernels and reads one 64-byte status record per restart cycle.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include <cooperative_groups.h>

#include "comm.cuh"
#include "common.cuh"
#include "host_math.h"
#include "vector_ops.cuh"

namespace cg = cooperative_groups;

namespace tn {
long long plan_dim(const tn_effh_plan* P);
bool plan_is_rows(const tn_effh_plan* P);
long long plan_slice_offset(const tn_effh_plan* P);
long long plan_slice_len(const tn_effh_plan* P);

struct LanczosState {
  double alpha[kMaxNcv];
  double beta[kMaxNcv];   // beta[j] couples v_j and v_{j+1}
  double h[kMaxNcv + 2];  // CGS pass 1 coefficients; h[j+1] = w.w of step j (DGKS test)
  double h2[kMaxNcv + 2]; // CGS pass 2 coefficients; sharded path: h2[j+1] = w'.w' after the first pass
  double lock_h[kMaxNcv]; // coefficients against locked (deflated) vectors, tn_lanczos_generic
  double u[kMaxNcv];      // Ritz vector in the Krylov basis
  double nrm2, inv_beta;
  double scale2[2];              // power-of-two pre-scaling of the start vector (its square must not overflow)
  unsigned long long scale_slot;
  int need2, pad2;        // DGKS: second Gram-Schmidt pass needed for the current step
  // status record copied to the host once per cycle
  double theta, resid, lambda, s_restart;
  int converged, breakdown, m_eff, ql_fail;
};

__global__ void lanczos_init_kernel(LanczosState* st) {
  st->breakdown = 0;
  st->m_eff = 0;
  st->converged = 0;
  st->ql_fail = 0;
  st->inv_beta = 0.0;
  const double n2 = st->nrm2;
  st->inv_beta = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
}

// DGKS test after the first Gram-Schmidt pass of step j: |w'|^2 = |w|^2 - sum h_i^2 (Pythagoras); a second pass is needed
// only when the first one removed more than half of |w|^2 (ARPACK applies the same test)
__global__ void lanczos_dgks_kernel(LanczosState* st, int j) {
  const double ww = st->h[j + 1];
  double s = 0.0;
  for (int i = 0; i <= j; ++i) s += st->h[i] * st->h[i];
  const int need = !((ww - s) > 0.5 * ww);
  st->need2 = need;
  if (!need) {
    for (int i = 0; i <= j; ++i) st->h2[i] = 0.0;
    st->h2[j + 1] = ww - s;   // |w'|^2 by Pythagoras (read by the sliced path in place of the measured w'.w')
  }
}

// after the two CGS passes and the norm of step j
__global__ void lanczos_finish_step_kernel(LanczosState* st, int j) {
  const double a = st->h[j] + st->h2[j];
  st->alpha[j] = a;
  const double bj = sqrt(st->nrm2);
  st->beta[j] = bj;
  double scale = fabs(a);
  if (j > 0) scale = fmax(scale, fabs(st->beta[j - 1]));
  scale = fmax(scale, 1e-300);
  if (st->breakdown || bj <= 1e-14 * scale) {
    if (!st->breakdown) {
      st->breakdown = 1;
      st->m_eff = j + 1;
      st->beta[j] = 0.0;
    }
    st->inv_beta = 0.0;  // v_{j+1} = 0: later steps of this cycle are inert
  } else {
    st->inv_beta = 1.0 / bj;
  }
}

// Row-sharded basis (every rank holds a slice of each Krylov vector): step j after BOTH Gram-Schmidt passes were applied.
// h2[j+1] holds the all-reduced w'.w' measured after the first pass, so |w''|^2 = w'.w' - sum h2_i^2 (Pythagoras) costs no
// third collective.  It is accurate when the second pass removed little; if it removed more than half of |w'|^2 the vector
// is numerically inside span(V) and the step is a breakdown (the DGKS rule ARPACK applies after its second pass).
__global__ void lanczos_finish_sharded_kernel(LanczosState* st, int j) {
  const double a = st->h[j] + st->h2[j];
  st->alpha[j] = a;
  const double w1 = st->h2[j + 1];
  double s = 0.0;
  for (int i = 0; i <= j; ++i) s += st->h2[i] * st->h2[i];
  double n2 = w1 - s;
  const bool inside = !(n2 > 0.5 * w1);
  if (n2 < 0.0) n2 = 0.0;
  st->nrm2 = n2;
  const double bj = sqrt(n2);
  st->beta[j] = bj;
  double scale = fabs(a);
  if (j > 0) scale = fmax(scale, fabs(st->beta[j - 1]));
  scale = fmax(scale, 1e-300);
  if (st->breakdown || inside || bj <= 1e-14 * scale) {
    if (!st->breakdown) {
      st->breakdown = 1;
      st->m_eff = j + 1;
      st->beta[j] = 0.0;
    }
    st->inv_beta = 0.0;
  } else {
    st->inv_beta = 1.0 / bj;
  }
}

// projected eigenproblem of dimension m (or m_eff after a breakdown); selects the Ritz value maximising |1 - tau*theta|.
// |1 - tau*theta| is convex in theta, so the wanted Ritz value is the smallest or the largest eigenvalue of T: both are found by
// 16-way multisection on Sturm counts (lanes 0-15 / 16-31, division-free minor recurrence), the eigenvector of the chosen one by
// the twisted factorisation (forward pivots on lane 0, backward pivots on lane 1) -- ~10 us instead of the ~160 us of the full
// QL, which stays as the fallback when the residual |T u - theta u| of the fast path is not at rounding level.
__global__ void __launch_bounds__(32) lanczos_ritz_kernel(LanczosState* st, int m_in, double tau, double tol, int force_ql) {
  constexpr int LDZ = kMaxNcv + 1;  // odd pitch: conflict-free row access
  constexpr unsigned kFull = 0xffffffffu;
  __shared__ double z[kMaxNcv * LDZ];
  __shared__ double sd[kMaxNcv], se[kMaxNcv], se2[kMaxNcv], spf[kMaxNcv], spb[kMaxNcv], su[kMaxNcv];
  const int lane = threadIdx.x;
  const int m = st->breakdown ? st->m_eff : m_in;
  const double beta_last = st->breakdown ? 0.0 : st->beta[m - 1];
  const double eps23 = 3.666852862501036e-11;  // eps^(2/3), ARPACK's floor on |lambda|
  if (!force_ql && m >= 2) {
    // Gershgorin disc, matrix normalised to [-1, 1]
    double glo = 1e300, ghi = -1e300;
    for (int i = lane; i < m; i += 32) {
      const double r = (i > 0 ? fabs(st->beta[i - 1]) : 0.0) + (i + 1 < m ? fabs(st->beta[i]) : 0.0);
      glo = fmin(glo, st->alpha[i] - r);
      ghi = fmax(ghi, st->alpha[i] + r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      glo = fmin(glo, __shfl_xor_sync(kFull, glo, o));
      ghi = fmax(ghi, __shfl_xor_sync(kFull, ghi, o));
    }
    const double centre = 0.5 * (glo + ghi), radius = fmax(0.5 * (ghi - glo), 1e-300), inv_r = 1.0 / radius;
    for (int i = lane; i < m; i += 32) {
      sd[i] = (st->alpha[i] - centre) * inv_r;
      const double en = (i + 1 < m) ? st->beta[i] * inv_r : 0.0;
      se[i] = en;
      se2[i] = en * en;
    }
    __syncwarp();
    // bracket of the NORMALISED spectrum (centre / radius above are rounded: the disc of the scaled matrix is taken anew)
    double lo = 1e300, hi = -1e300;
    for (int i = lane; i < m; i += 32) {
      const double r = (i > 0 ? fabs(se[i - 1]) : 0.0) + fabs(se[i]);
      lo = fmin(lo, sd[i] - r);
      hi = fmax(hi, sd[i] + r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(kFull, lo, o));
      hi = fmax(hi, __shfl_xor_sync(kFull, hi, o));
    }
    lo -= 0x1p-40;
    hi += 0x1p-40;
    const int half = lane >> 4, idx = lane & 15, k = half ? m - 1 : 0;
    for (int round = 0; round < 48; ++round) {
      const double x = multisect_point(lo, hi, 16, idx);
      const int c = tridiag_count_below(m, sd, se2, x);
      const unsigned mask = __ballot_sync(kFull, c >= k + 1);
      const unsigned mine = (mask >> (16 * half)) & 0xffffu;
      multisect_shrink(&lo, &hi, 16, mine ? __ffs((int)mine) - 1 : 16);
      const bool done = !((hi - lo) > 2.5e-16) || !(multisect_point(lo, hi, 16, 0) > lo);
      if (__all_sync(kFull, done)) break;
    }
    const double x_mine = 0.5 * (lo + hi);
    const double x_min = __shfl_sync(kFull, x_mine, 0), x_max = __shfl_sync(kFull, x_mine, 16);
    const double th_min = centre + radius * x_min, th_max = centre + radius * x_max;
    const bool pick_max = fabs(1.0 - tau * th_max) > fabs(1.0 - tau * th_min);
    const double xs = pick_max ? x_max : x_min, theta = pick_max ? th_max : th_min;
    if (lane == 0) twisted_pivots(m, sd, se, xs, +1, 1e-280, spf);
    if (lane == 1) twisted_pivots(m, sd, se, xs, -1, 1e-280, spb);
    __syncwarp();
    if (lane == 0) twisted_vector(m, sd, se, xs, spf, spb, su);
    __syncwarp();
    double n2 = 0.0;
    for (int i = lane; i < m; i += 32) n2 += su[i] * su[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(kFull, n2, o);
    const double inv_n = rsqrt(n2);
    double rmax = 0.0;
    for (int i = lane; i < m; i += 32) {
      const double r = (sd[i] - xs) * su[i] + (i > 0 ? se[i - 1] * su[i - 1] : 0.0) + (i + 1 < m ? se[i] * su[i + 1] : 0.0);
      rmax = fmax(rmax, fabs(r) * inv_n);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(kFull, rmax, o));
    const bool ok = n2 > 0.0 && n2 < 1e300 && rmax <= 2e-14;   // NaN fails the comparison too
    if (ok) {
      for (int i = lane; i < kMaxNcv; i += 32) st->u[i] = i < m ? su[i] * inv_n : 0.0;
      if (lane == 0) {
        const double s = beta_last * su[m - 1] * inv_n;
        st->ql_fail = 0;
        st->theta = theta;
        st->lambda = 1.0 - tau * theta;
        st->s_restart = s;
        st->resid = fabs(s);
        st->converged = (st->breakdown || fabs(tau) * fabs(s) <= tol * fmax(eps23, fabs(st->lambda))) ? 1 : 0;
        st->m_eff = m;
      }
      return;
    }
    __syncwarp();
  }
  // full QL: every lane runs the scalar recurrences on private copies of (d, e); lane k owns rows k, k+32 of z
  double d[kMaxNcv], e[kMaxNcv];
  for (int i = 0; i < m; ++i) {
    d[i] = st->alpha[i];
    e[i] = st->beta[i];
  }
  for (int i = lane; i < m; i += 32)
    for (int k = 0; k < m; ++k) z[i * LDZ + k] = (i == k) ? 1.0 : 0.0;
  __syncwarp();
  const int fail = tridiag_ql_rows(m, d, e, z, LDZ, lane, 32);
  __syncwarp();
  int best = 0;
  for (int k = 1; k < m; ++k)
    if (fabs(1.0 - tau * d[k]) > fabs(1.0 - tau * d[best])) best = k;
  if (lane != 0) return;
  double nrm = 0.0;
  for (int i = 0; i < m; ++i) nrm += z[i * LDZ + best] * z[i * LDZ + best];
  nrm = sqrt(nrm);
  for (int i = 0; i < kMaxNcv; ++i) st->u[i] = i < m ? z[i * LDZ + best] / nrm : 0.0;
  const double s = beta_last * st->u[m - 1];
  st->ql_fail = fail;
  st->theta = d[best];
  st->lambda = 1.0 - tau * d[best];
  st->s_restart = s;
  st->resid = fabs(s);
  st->converged = (st->breakdown || fabs(tau) * fabs(s) <= tol * fmax(eps23, fabs(st->lambda))) ? 1 : 0;
  st->m_eff = m;
}

// thick restart with one kept Ritz pair: basis (y, r_hat), T = [[theta, s], [s, .]]
__global__ void lanczos_restart_kernel(LanczosState* st) {
  st->alpha[0] = st->theta;
  st->beta[0] = st->s_restart;
  st->breakdown = 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Fused re-orthogonalisation step for vectors that live in L2 (n <= kFusedMaxN): one cooperative launch replaces the
// seven launches of the streaming path.  Every CTA owns a contiguous slice of the vector, keeps its slice of w in shared
// memory across the phases, and the grid meets at three grid-wide barriers:
//   phase 1  h  = V^T w            (warp-per-vector dots over the slice, warp-shuffle reduction)   -> sync
//   phase 2  w -= V h ; h2 = V^T w                                                                 -> sync
//   phase 3  w -= V h2 ; |w|^2                                                                     -> sync
//   phase 4  alpha_j, beta_j, breakdown test (same arithmetic as lanczos_finish_step_kernel); w *= 1/beta, written back
// Partial sums are combined in CTA order by every CTA, so the result is bit-reproducible and identical in all CTAs.
constexpr int kFusedThreads = 256;
constexpr long long kFusedMaxN = 1LL << 19;       // 4 MiB vectors: 22 of them stay L2 resident

__device__ __forceinline__ double warp_sum_l(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kFusedThreads) lanczos_orth_fused_kernel(double* __restrict__ V, long long ldv, int j, long long n,
                                                                           LanczosState* st, double* __restrict__ partial,
                                                                           int slice_cap, int cache_v, int bulk) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double dyn[];
  double* ws = dyn;              // this CTA's slice of w (slice_cap doubles)
  double* vs = dyn + slice_cap;  // cache_v: this CTA's slice of v_0..v_j, row pitch slice_cap
  __shared__ double hs[kMaxNcv + 1];
  __shared__ double red[kFusedThreads / 32];
  const int nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nvec = j + 1;
  // slices of slice_cap elements (even): CTA b owns [b * slice_cap, ...); with bulk staging the slices are 16-byte granular
  const long long chunk = bulk ? (long long)slice_cap : (n + nb - 1) / nb;
  const long long e0 = min((long long)b * chunk, n);
  const int len = (int)(min(e0 + chunk, n) - e0);
  double* w = V + (long long)(j + 1) * ldv + e0;
  const double* Vg = V + e0;
  if (bulk) {
    // the CTA's slice of w and of v_0..v_j: one 1-D bulk copy per vector (cp.async.bulk, UBLKCP), all in flight at once and
    // signalled on one mbarrier -- a plain load/store loop serialises ~20 L2 round trips here (40 % of the kernel was spent
    // waiting on them, profiles/r02_small_chi.md)
    __shared__ uint64_t bar;
    const unsigned bytes = (unsigned)(len * sizeof(double));   // len is even: multiple of 16
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && len > 0) {
      const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes * (unsigned)(1 + (cache_v ? nvec : 0))) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       (unsigned)__cvta_generic_to_shared(ws)), "l"(w), "r"(bytes), "r"(bar_a) : "memory");
      if (cache_v)
        for (int i = 0; i < nvec; ++i)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           (unsigned)__cvta_generic_to_shared(vs + (size_t)i * slice_cap)), "l"(Vg + (long long)i * ldv), "r"(bytes), "r"(bar_a) : "memory");
    }
    if (len > 0) {
      asm volatile(
          "{\n.reg .pred p;\nLZ_WAIT_%=:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
          "@!p bra LZ_WAIT_%=;\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
    }
  } else {
    for (int e = tid; e < len; e += kFusedThreads) ws[e] = w[e];
    if (cache_v) {  // one coalesced pass over the basis slice; all later phases run out of shared memory
      for (int i = 0; i < nvec; ++i) {
        const double* v = Vg + (long long)i * ldv;
        for (int e = tid; e < len; e += kFusedThreads) vs[i * slice_cap + e] = v[e];
      }
    }
  }
  __syncthreads();
  const double* Vs = cache_v ? vs : Vg;
  const long long vpitch = cache_v ? (long long)slice_cap : ldv;

  auto dots = [&](double* out /* nb x (kMaxNcv+1) */) {
    for (int i = warp; i < nvec; i += kFusedThreads / 32) {
      const double* v = Vs + (long long)i * vpitch;
      double acc0 = 0.0, acc1 = 0.0;
      int e = lane;
      for (; e + 32 < len; e += 64) {
        acc0 += v[e] * ws[e];
        acc1 += v[e + 32] * ws[e + 32];
      }
      if (e < len) acc0 += v[e] * ws[e];
      const double acc = warp_sum_l(acc0 + acc1);
      if (lane == 0) out[(long long)b * (kMaxNcv + 1) + i] = acc;
    }
  };
  auto gather = [&](const double* in) {  // hs[i] = sum over CTAs: lane-strided partial sums + fixed shuffle tree (same in all CTAs)
    __syncthreads();
    for (int i = warp; i < nvec; i += kFusedThreads / 32) {
      double s = 0.0;
      for (int c = lane; c < nb; c += 32) s += in[(long long)c * (kMaxNcv + 1) + i];
      s = warp_sum_l(s);
      if (lane == 0) hs[i] = s;
    }
    __syncthreads();
  };
  auto update = [&]() {  // ws -= sum_i hs[i] * V_i
    for (int e = tid; e < len; e += kFusedThreads) {
      double v0 = ws[e], v1 = 0.0;
      int i = 0;
      for (; i + 1 < nvec; i += 2) {
        v0 -= hs[i] * Vs[(long long)i * vpitch + e];
        v1 -= hs[i + 1] * Vs[(long long)(i + 1) * vpitch + e];
      }
      if (i < nvec) v0 -= hs[i] * Vs[(long long)i * vpitch + e];
      ws[e] = v0 + v1;
    }
    __syncthreads();
  };
  double* p1 = partial;
  double* p2 = partial + (long long)nb * (kMaxNcv + 1);
  double* p3 = p2 + (long long)nb * (kMaxNcv + 1);
  auto sumsq = [&](double* out) {  // out[b] = sum of ws^2 over this CTA's slice
    double acc = 0.0;
    for (int e = tid; e < len; e += kFusedThreads) acc += ws[e] * ws[e];
    acc = warp_sum_l(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < kFusedThreads / 32; ++k) s += red[k];
      out[b] = s;
    }
  };

  dots(p1);
  grid.sync();
  gather(p1);
  const double h_j = hs[j];
  update();
  dots(p2);
  sumsq(p3);   // |w'|^2 after the first pass, in the same grid phase as the second-pass coefficients
  grid.sync();
  gather(p2);
  const double h2_j = hs[j];
  double w1 = 0.0;
  for (int c = lane; c < nb; c += 32) w1 += p3[c];
  w1 = warp_sum_l(w1);  // every warp of every CTA computes the same value
  double s2 = 0.0;
  for (int i = 0; i < nvec; ++i) s2 += hs[i] * hs[i];
  update();
  // |w''|^2 = |w'|^2 - sum h2_i^2 (Pythagoras): accurate while the second pass removes little; if it removed more than half of
  // |w'|^2 the vector is numerically inside span(V) and the step is a breakdown (the DGKS rule ARPACK applies after its second
  // pass).  Two grid-wide barriers per step instead of four.
  double nrm2 = w1 - s2;
  const bool inside = !(nrm2 > 0.5 * w1);
  if (nrm2 < 0.0) nrm2 = 0.0;
  // bookkeeping, identical in every thread of every CTA.  beta[j-1] was written by an earlier launch; `breakdown` may be
  // rewritten by CTA 0 below while other CTAs still read it, which is benign: it only ever flips 0 -> 1 in a step whose own
  // test (`inside` / tiny beta, the same data in every CTA) says "broken" as well.
  const double a = h_j + h2_j;
  const double bj = sqrt(nrm2);
  double scale = fabs(a);
  if (j > 0) scale = fmax(scale, fabs(st->beta[j - 1]));
  scale = fmax(scale, 1e-300);
  const int was_broken = *((volatile int*)&st->breakdown);
  const bool broken = was_broken || inside || bj <= 1e-14 * scale;
  const double inv = broken ? 0.0 : 1.0 / bj;
  for (int e = tid; e < len; e += kFusedThreads) w[e] = ws[e] * inv;
  if (b == 0 && tid == 0) {
    st->alpha[j] = a;
    st->beta[j] = bj;
    st->nrm2 = nrm2;
    if (broken) {
      if (!was_broken) {
        st->breakdown = 1;
        st->m_eff = j + 1;
        st->beta[j] = 0.0;
      }
      st->inv_beta = 0.0;
    } else {
      st->inv_beta = inv;
    }
  }
}

struct LanczosStatus {
  double theta, resid, lambda, s_restart;
  int converged, breakdown, m_eff, ql_fail;
};

// What the driver needs from an operator: y_loc = H x, where x is Krylov vector `x_loc` (this rank's slice, or the whole
// vector when nothing is sharded) and y_loc the matching slice of the result.
struct LanczosOp {
  tn_effh_plan* plan = nullptr;      // effective-Hamiltonian plan (full, term-sharded or row-sliced)
  tn_matvec_fn fn = nullptr;         // or a caller-supplied operator (tn_lanczos_generic)
  void* fn_user = nullptr;
  tn_allreduce_fn allreduce = nullptr;  // term-sharded plan, caller's collective
  void* allreduce_user = nullptr;
  tn_comm* comm = nullptr;           // in-library collectives
  bool rows = false;                 // row-sliced plan: gather the slices of x into `full`, apply the plan to it
  double* full = nullptr;            // world * n_pad doubles
  long long n = 0, n_pad = 0;

  int apply(const double* x_loc, double* y_loc, cudaStream_t stream) {
    if (fn) {
      int rc = fn(x_loc, y_loc, fn_user, stream);
      TN_REQUIRE(rc == 0, "tn_lanczos: matvec callback failed (%d)", rc);
      return TN_OK;
    }
    if (rows) {
      // the gathered vector arrives in the library's peer window (plain stores over NVLink) when the box allows it, else in `full`
      const double* gathered = nullptr;
      TN_CHECK(comm_allgather_window(comm, x_loc, n_pad, full, &gathered, stream));
      return tn_effh_matvec(plan, gathered, y_loc, 0.0, 1.0, stream);
    }
    TN_CHECK(tn_effh_matvec(plan, x_loc, y_loc, 0.0, 1.0, stream));
    if (comm) return comm_allreduce_sum(comm, y_loc, n, stream);
    if (allreduce) {
      int rc = allreduce(y_loc, n, allreduce_user, stream);
      TN_REQUIRE(rc == 0, "tn_lanczos_lm1: all-reduce callback failed (%d)", rc);
    }
    return TN_OK;
  }
};

static size_t lanczos_ws_bytes(long long n_loc_pad, int ncv, long long full_len) {
  const long long m = std::min<long long>(std::max(ncv, 2), kMaxNcv);
  const size_t ldv = (size_t)((n_loc_pad + 1) / 2 * 2);
  return align_up(sizeof(double) * ldv * (size_t)(m + 2)) +
         align_up(sizeof(double) * std::max<size_t>((size_t)(m + 2) * dot_chunks(n_loc_pad), (size_t)3 * 2048 * (kMaxNcv + 1))) +
         align_up(sizeof(LanczosState)) + align_up(sizeof(unsigned)) + align_up(sizeof(double) * (size_t)full_len) + 1024;
}

// n: global dimension; n_loc: length of this rank's slice (== n when nothing is sliced); off: its offset in the vector
static int lanczos_core(LanczosOp& op, long long n, long long n_loc, long long off, double tau, const double* v0, double tol,
                        int ncv, int max_restarts, const double* locked, int n_locked, long long ld_locked,
                        double* lambda_out, double* vec_out, int* n_matvec_out, double* resid_out, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream) {
  const bool sliced = op.rows;
  tn_comm* comm = op.comm;
  const long long n_store = sliced ? op.n_pad : n;  // stored length of a basis vector (the pad is zero and never written)
  const long long full_len = sliced ? op.n_pad * comm->world : 0;
  if (workspace_bytes < lanczos_ws_bytes(n_store, ncv, full_len)) {
    set_error("tn_lanczos: workspace %zu < %zu bytes", workspace_bytes, lanczos_ws_bytes(n_store, ncv, full_len));
    return TN_ERR_WORKSPACE;
  }
  const int m = (int)std::min<long long>(std::max(ncv, 2), std::min<long long>(n - n_locked, kMaxNcv));
  TN_REQUIRE(m >= 1, "tn_lanczos: nothing left to solve (n = %lld, locked = %d)", n, n_locked);
  const long long ldv = (n_store + 1) / 2 * 2;
  Carver cw(workspace, workspace_bytes);
  double* V = cw.take<double>((size_t)ldv * (m + 2));
  double* partial = cw.take<double>(std::max<size_t>((size_t)(m + 2) * dot_chunks(n_store), (size_t)3 * 2048 * (kMaxNcv + 1)));
  LanczosState* st = cw.take<LanczosState>(1);
  unsigned* counter = cw.take<unsigned>(1);
  double* full = sliced ? cw.take<double>((size_t)full_len) : nullptr;
  TN_REQUIRE(V && partial && st && counter && (!sliced || full), "tn_lanczos: workspace carve failed");
  op.full = full;
  double* ytmp = V + (size_t)ldv * (m + 1);
  auto vec = [&](int j) { return V + (size_t)ldv * j; };
  // dot products of slices are partial sums: one small all-reduce makes them global (and identical on every rank)
  auto reduce = [&](double* dev, int count) -> int { return sliced ? comm_allreduce_sum(comm, dev, count, stream) : TN_OK; };
  auto project_locked = [&](double* w) -> int {  // w -= Lk (Lk^T w), twice (locked vectors are orthonormal)
    for (int pass = 0; pass < 2 && n_locked > 0; ++pass) {
      TN_CHECK(launch_multidot(locked, ld_locked, n_locked, w, n_loc, st->lock_h, partial, counter, stream));
      TN_CHECK(launch_multi_axpy(w, locked, ld_locked, n_locked, st->lock_h, n_loc, stream));
    }
    return TN_OK;
  };

  TN_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
  TN_CUDA(cudaMemsetAsync(st, 0, sizeof(LanczosState), stream));
  if (sliced) TN_CUDA(cudaMemsetAsync(V, 0, sizeof(double) * (size_t)ldv * (m + 2), stream));  // the pads travel through all-gathers
  // v_0 = v0 / |v0|
  TN_CUDA(cudaMemcpyAsync(vec(0), v0 + off, sizeof(double) * n_loc, cudaMemcpyDeviceToDevice, stream));
  // an un-normalised centre tensor of a long random chain can carry a norm of 1e300: scale by a power of two before squaring
  // (the scale is taken from the full vector, which every rank holds, so sliced ranks agree on it)
  TN_CHECK(launch_pow2_scale(v0, n, st->scale2, &st->scale_slot, stream));
  TN_CHECK(launch_scale_dev(vec(0), st->scale2, n_loc, stream));
  TN_CHECK(project_locked(vec(0)));
  TN_CHECK(launch_multidot(vec(0), ldv, 1, vec(0), n_loc, &st->nrm2, partial, counter, stream));
  TN_CHECK(reduce(&st->nrm2, 1));
  lanczos_init_kernel<<<1, 1, 0, stream>>>(st);
  TN_LAUNCHED();
  TN_CHECK(launch_scale_dev(vec(0), &st->inv_beta, n_loc, stream));

  // fused cooperative re-orthogonalisation when the vectors are L2 resident and a slice fits the shared-memory buffer
  int fused_grid = 0, fused_slice = 0, fused_cache = 0, fused_bulk = 0;
  size_t fused_smem = 0;
  if (!sliced && n <= kFusedMaxN && !getenv("TNALG_NO_FUSED_ORTH")) {
    const int sms = sm_count();
    // preferred: one CTA per SM with the CTA's slice of the whole basis cached in shared memory
    const int grid1 = (int)std::max<long long>(1, std::min<long long>(sms, (n + 511) / 512));
    const int slice1 = (int)(((n + grid1 - 1) / grid1 + 1) / 2 * 2);
    const size_t smem1 = sizeof(double) * (size_t)slice1 * (size_t)(m + 2);
    if (smem1 <= 200 * 1024) {
      fused_grid = grid1; fused_slice = slice1; fused_cache = 1; fused_smem = smem1;
      // bulk-copy staging needs 16-byte granular slices: even n (ldv is even, the workspace 256-byte aligned)
      fused_bulk = (n % 2 == 0 && (reinterpret_cast<uintptr_t>(V) & 15) == 0 && !getenv("TNALG_NO_BULK_ORTH")) ? 1 : 0;
    }  // larger vectors: the streaming kernels (more CTAs in flight) are faster than an uncached fused pass
    if (fused_grid > 0) {
      static size_t configured = 0;
      if (fused_smem > configured) {
        TN_CUDA(cudaFuncSetAttribute(lanczos_orth_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        configured = 200 * 1024;
      }
      int per_sm = 0;
      TN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lanczos_orth_fused_kernel, kFusedThreads, fused_smem));
      if (per_sm * sms < fused_grid) fused_grid = 0;  // cannot be co-resident: use the streaming path
    }
  }
  // sliced basis: with the peer window the second Gram-Schmidt pass is predicated on the device (DGKS)
  const bool peer_pred = sliced && m + 2 <= kPeerArMax && !getenv("TNALG_NO_DGKS_SKIP") && comm_peer_available(comm, stream);
  static const int ritz_force_ql = getenv("TNALG_RITZ_QL") ? 1 : 0;   // A/B timing of the projected eigenproblem
  int n_matvec = 0;
  int j0 = 0;
  LanczosStatus hs{};
  int status = TN_ERR_NOCONV;
  for (int cycle = 0; cycle < max_restarts; ++cycle) {
    for (int j = j0; j < m; ++j) {
      double* w = vec(j + 1);
      TN_CHECK(op.apply(vec(j), w, stream));
      ++n_matvec;
      TN_CHECK(project_locked(w));
      if (sliced && peer_pred) {
        // sliced basis, collectives in the peer window: the DGKS flag derives from all-reduced numbers (identical on every
        // rank), so the second Gram-Schmidt pass AND its all-reduce are skipped on the device when the first pass removed
        // less than half of |w|^2; |w'|^2 = |w|^2 - sum h^2 then stands in for the measured norm (h2[j+1])
        TN_CHECK(launch_multidot(V, ldv, j + 2, w, n_loc, st->h, partial, counter, stream));
        TN_CHECK(reduce(st->h, j + 2));
        lanczos_dgks_kernel<<<1, 1, 0, stream>>>(st, j);
        TN_LAUNCHED();
        TN_CHECK(launch_multi_axpy(w, V, ldv, j + 1, st->h, n_loc, stream));
        TN_CHECK(launch_multidot(V, ldv, j + 2, w, n_loc, st->h2, partial, counter, stream, &st->need2));
        TN_CHECK(comm_allreduce_sum(comm, st->h2, j + 2, stream, &st->need2));
        TN_CHECK(launch_multi_axpy(w, V, ldv, j + 1, st->h2, n_loc, stream, &st->need2));
        lanczos_finish_sharded_kernel<<<1, 1, 0, stream>>>(st, j);
        TN_LAUNCHED();
        TN_CHECK(launch_scale_dev(w, &st->inv_beta, n_loc, stream));
      } else if (sliced) {
        // sliced basis: full CGS2 with two small all-reduces per step (h and w.w; h2 and w'.w')
        TN_CHECK(launch_multidot(V, ldv, j + 2, w, n_loc, st->h, partial, counter, stream));
        TN_CHECK(reduce(st->h, j + 2));
        TN_CHECK(launch_multi_axpy(w, V, ldv, j + 1, st->h, n_loc, stream));
        TN_CHECK(launch_multidot(V, ldv, j + 2, w, n_loc, st->h2, partial, counter, stream));
        TN_CHECK(reduce(st->h2, j + 2));
        TN_CHECK(launch_multi_axpy(w, V, ldv, j + 1, st->h2, n_loc, stream));
        lanczos_finish_sharded_kernel<<<1, 1, 0, stream>>>(st, j);
        TN_LAUNCHED();
        TN_CHECK(launch_scale_dev(w, &st->inv_beta, n_loc, stream));
      } else if (fused_grid > 0) {
        // one cooperative launch: CGS2, norm, scale and the alpha/beta bookkeeping of step j
        long long ldv_arg = ldv, n_arg = n;
        int j_arg = j;
        int bulk_arg = fused_bulk;
        void* args[] = {(void*)&V, (void*)&ldv_arg, (void*)&j_arg, (void*)&n_arg, (void*)&st, (void*)&partial,
                        (void*)&fused_slice, (void*)&fused_cache, (void*)&bulk_arg};
        TN_CUDA(cudaLaunchCooperativeKernel((const void*)lanczos_orth_fused_kernel, dim3(fused_grid), dim3(kFusedThreads), args,
                                            fused_smem, stream));
        TN_LAUNCHED();
      } else {
        // streaming path (HBM-bound kernels): CGS2 against v_0..v_j, norm, bookkeeping, scale
        // pass 1 also yields w.w (w = v_{j+1} is the vector after v_j in the workspace) for the DGKS test
        TN_CHECK(launch_multidot(V, ldv, j + 2, w, n, st->h, partial, counter, stream));
        lanczos_dgks_kernel<<<1, 1, 0, stream>>>(st, j);
        TN_LAUNCHED();
        TN_CHECK(launch_multi_axpy(w, V, ldv, j + 1, st->h, n, stream));
        TN_CHECK(launch_multidot(V, ldv, j + 1, w, n, st->h2, partial, counter, stream, &st->need2));
        TN_CHECK(launch_multi_axpy(w, V, ldv, j + 1, st->h2, n, stream, &st->need2));
        TN_CHECK(launch_multidot(w, ldv, 1, w, n, &st->nrm2, partial, counter, stream));
        lanczos_finish_step_kernel<<<1, 1, 0, stream>>>(st, j);
        TN_LAUNCHED();
        TN_CHECK(launch_scale_dev(w, &st->inv_beta, n, stream));
      }
    }
    lanczos_ritz_kernel<<<1, 32, 0, stream>>>(st, m, tau, tol, ritz_force_ql);
    TN_LAUNCHED();
    TN_CUDA(cudaMemcpyAsync(&hs, &st->theta, sizeof(LanczosStatus), cudaMemcpyDeviceToHost, stream));
    TN_CUDA(cudaStreamSynchronize(stream));
    if (hs.ql_fail) {
      // nothing has been written to the outputs: a hard error, not the soft "iteration limit" status
      set_error("tn_lanczos: the projected tridiagonal eigenproblem did not converge (NaN/Inf in the operator or the start vector?)");
      return TN_ERR_NUMERIC;
    }
    // Ritz vector y = V u
    TN_CHECK(launch_combine(ytmp, V, ldv, hs.m_eff, st->u, n_loc, stream));
    if (hs.converged || m >= n - n_locked) {
      status = TN_OK;
      break;
    }
    if (cycle + 1 == max_restarts) break;
    // thick restart: v_0 = y, v_1 = residual direction (old v_m), T[0,0] = theta, T[0,1] = s
    TN_CUDA(cudaMemcpyAsync(vec(0), ytmp, sizeof(double) * n_loc, cudaMemcpyDeviceToDevice, stream));
    TN_CUDA(cudaMemcpyAsync(vec(1), vec(m), sizeof(double) * n_loc, cudaMemcpyDeviceToDevice, stream));
    lanczos_restart_kernel<<<1, 1, 0, stream>>>(st);
    TN_LAUNCHED();
    j0 = 1;
  }
  // normalise and return
  TN_CHECK(launch_multidot(ytmp, ldv, 1, ytmp, n_loc, &st->nrm2, partial, counter, stream));
  TN_CHECK(reduce(&st->nrm2, 1));
  lanczos_init_kernel<<<1, 1, 0, stream>>>(st);
  TN_LAUNCHED();
  TN_CHECK(launch_scale_dev(ytmp, &st->inv_beta, n_loc, stream));
  if (sliced) {  // every rank receives the whole eigenvector (bit-identical on all ranks)
    const double* gathered = nullptr;
    TN_CHECK(comm_allgather_window(comm, ytmp, op.n_pad, full, &gathered, stream));
    TN_CUDA(cudaMemcpyAsync(vec_out, gathered, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
  } else {
    TN_CUDA(cudaMemcpyAsync(vec_out, ytmp, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
  }
  TN_CUDA(cudaStreamSynchronize(stream));
  if (lambda_out) *lambda_out = hs.lambda;
  if (resid_out) *resid_out = hs.resid;
  if (n_matvec_out) *n_matvec_out = n_matvec;
  if (status == TN_ERR_NOCONV) set_error("tn_lanczos: not converged after %d restart cycles (residual %.3e)", max_restarts, hs.resid);
  return status;
}

}  // namespace tn

using namespace tn;

// slice geometry of a row-sliced plan inside a communicator: rows_per = ceil(a / world), rank r owns rows [r*rows_per, ...)
static int sliced_geometry(tn_effh_plan* plan, tn_comm* comm, long long* n_pad) {
  const long long n = plan_dim(plan), n_out = plan_slice_len(plan), off = plan_slice_offset(plan);
  int rb = 0, rc = 0;
  tn_effh_plan_rows(plan, &rb, &rc);
  TN_REQUIRE(comm && comm->world >= 1, "tn_lanczos_lm1: a row-sliced plan needs a communicator");
  const long long per_row = n_out / rc;                 // d * b
  const long long a = n / per_row;
  const long long rows_per = (a + comm->world - 1) / comm->world;
  TN_REQUIRE((long long)rb == rows_per * comm->rank && rc == std::min<long long>(rows_per, a - rb) && off == rb * per_row,
             "tn_lanczos_lm1: plan rows [%d, %d) do not match rank %d of %d (rows per rank %lld)", rb, rb + rc, comm->rank,
             comm->world, rows_per);
  *n_pad = rows_per * per_row;
  return TN_OK;
}

extern "C" size_t tn_lanczos_workspace_bytes(long long n, int ncv) { return lanczos_ws_bytes(n, ncv, 0); }

extern "C" size_t tn_lanczos_workspace_bytes_sharded(const tn_effh_plan* plan, const tn_comm* comm, int ncv) {
  if (!plan) return 0;
  if (!plan_is_rows(plan) || !comm) return lanczos_ws_bytes(plan_dim(plan), ncv, 0);
  long long n_pad = 0;
  if (sliced_geometry(const_cast<tn_effh_plan*>(plan), const_cast<tn_comm*>(comm), &n_pad) != TN_OK) return 0;
  return lanczos_ws_bytes(n_pad, ncv, n_pad * comm->world);
}

extern "C" int tn_lanczos_lm1(tn_effh_plan* plan, double tau, const double* v0, double tol, int ncv, int max_restarts,
                              double* lambda_out, double* vec_out, int* n_matvec_out, double* resid_out,
                              tn_allreduce_fn allreduce, void* allreduce_user, tn_comm* comm, void* workspace,
                              size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(plan && v0 && vec_out, "tn_lanczos_lm1: null argument");
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_lanczos_lm1: workspace must be 256-byte aligned");
  TN_REQUIRE(tol >= 0 && max_restarts >= 1, "tn_lanczos_lm1: bad tol/max_restarts");
  TN_REQUIRE(!(allreduce && comm), "tn_lanczos_lm1: pass either an all-reduce callback or a communicator, not both");
  LanczosOp op;
  op.plan = plan; op.allreduce = allreduce; op.allreduce_user = allreduce_user; op.comm = comm;
  op.n = plan_dim(plan);
  long long n_loc = op.n, off = 0;
  if (plan_is_rows(plan)) {
    TN_CHECK(sliced_geometry(plan, comm, &op.n_pad));
    op.rows = true;
    n_loc = plan_slice_len(plan);
    off = plan_slice_offset(plan);
  }
  return lanczos_core(op, op.n, n_loc, off, tau, v0, tol, ncv, max_restarts, nullptr, 0, 0, lambda_out, vec_out, n_matvec_out,
                      resid_out, workspace, workspace_bytes, stream);
}

extern "C" int tn_lanczos_generic(tn_matvec_fn matvec, void* user, long long n, double tau, const double* v0, double tol, int ncv,
                                  int max_restarts, const double* locked, int n_locked, long long ld_locked, double* lambda_out,
                                  double* vec_out, int* n_matvec_out, double* resid_out, void* workspace, size_t workspace_bytes,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(matvec && v0 && vec_out && n > 0, "tn_lanczos_generic: null argument");
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_lanczos_generic: workspace must be 256-byte aligned");
  TN_REQUIRE(tol >= 0 && max_restarts >= 1, "tn_lanczos_generic: bad tol/max_restarts");
  TN_REQUIRE(n_locked >= 0 && n_locked <= kMaxNcv && (n_locked == 0 || (locked && ld_locked >= n)), "tn_lanczos_generic: bad locked vectors");
  LanczosOp op;
  op.fn = matvec; op.fn_user = user; op.n = n;
  return lanczos_core(op, n, n, 0, tau, v0, tol, ncv, max_restarts, locked, n_locked, ld_locked, lambda_out, vec_out,
                      n_matvec_out, resid_out, workspace, workspace_bytes, stream);
}
