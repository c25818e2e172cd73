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
#include <vector>

#include "common.cuh"
#include "host_math.h"
#include "vector_ops.cuh"

namespace tn {
long long plan_dim(const tn_effh_plan* P);

struct LanczosState {
  double alpha[kMaxNcv];
  double beta[kMaxNcv];   // beta[j] couples v_j and v_{j+1}
  double h[kMaxNcv + 1];  // CGS pass 1 coefficients
  double h2[kMaxNcv + 1]; // CGS pass 2 coefficients
  double u[kMaxNcv];      // Ritz vector in the Krylov basis
  double nrm2, inv_beta;
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

// projected eigenproblem of dimension m (or m_eff after a breakdown); selects the Ritz value maximising |1 - tau*theta|
__global__ void lanczos_ritz_kernel(LanczosState* st, int m_in, double tau, double tol) {
  __shared__ double d[kMaxNcv], e[kMaxNcv], z[kMaxNcv * kMaxNcv];
  if (threadIdx.x != 0) return;
  const int m = st->breakdown ? st->m_eff : m_in;
  for (int i = 0; i < m; ++i) {
    d[i] = st->alpha[i];
    e[i] = st->beta[i];
    for (int k = 0; k < m; ++k) z[i * m + k] = (i == k) ? 1.0 : 0.0;
  }
  const double beta_last = st->breakdown ? 0.0 : st->beta[m - 1];
  st->ql_fail = tridiag_ql(m, d, e, z, m);
  int best = 0;
  for (int k = 1; k < m; ++k)
    if (fabs(1.0 - tau * d[k]) > fabs(1.0 - tau * d[best])) best = k;
  double nrm = 0.0;
  for (int i = 0; i < m; ++i) nrm += z[i * m + best] * z[i * m + best];
  nrm = sqrt(nrm);
  for (int i = 0; i < kMaxNcv; ++i) st->u[i] = i < m ? z[i * m + best] / nrm : 0.0;
  const double s = beta_last * st->u[m - 1];
  st->theta = d[best];
  st->lambda = 1.0 - tau * d[best];
  st->s_restart = s;
  st->resid = fabs(s);
  const double eps23 = 3.666852862501036e-11;  // eps^(2/3), ARPACK's floor on |lambda|
  st->converged = (st->breakdown || fabs(tau) * fabs(s) <= tol * fmax(eps23, fabs(st->lambda))) ? 1 : 0;
  st->m_eff = m;
}

// thick restart with one kept Ritz pair: basis (y, r_hat), T = [[theta, s], [s, .]]
__global__ void lanczos_restart_kernel(LanczosState* st) {
  st->alpha[0] = st->theta;
  st->beta[0] = st->s_restart;
  st->breakdown = 0;
}

struct LanczosStatus {
  double theta, resid, lambda, s_restart;
  int converged, breakdown, m_eff, ql_fail;
};

}  // namespace tn

using namespace tn;

extern "C" size_t tn_lanczos_workspace_bytes(long long n, int ncv) {
  const long long m = std::min<long long>(std::max(ncv, 2), std::min<long long>(n, kMaxNcv));
  const size_t ldv = (size_t)((n + 1) / 2 * 2);
  return align_up(sizeof(double) * ldv * (size_t)(m + 2)) + align_up(sizeof(double) * (size_t)(m + 1) * dot_chunks(n)) +
         align_up(sizeof(LanczosState)) + align_up(sizeof(unsigned)) + 1024;
}

extern "C" int tn_lanczos_lm1(tn_effh_plan* plan, double tau, const double* v0, double tol, int ncv, int max_restarts,
                              double* lambda_out, double* vec_out, int* n_matvec_out, double* resid_out,
                              tn_allreduce_fn allreduce, void* allreduce_user, void* workspace, size_t workspace_bytes,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(plan && v0 && vec_out, "tn_lanczos_lm1: null argument");
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_lanczos_lm1: workspace must be 256-byte aligned");
  const long long n = plan_dim(plan);
  if (workspace_bytes < tn_lanczos_workspace_bytes(n, ncv)) {
    set_error("tn_lanczos_lm1: workspace %zu < %zu bytes", workspace_bytes, tn_lanczos_workspace_bytes(n, ncv));
    return TN_ERR_WORKSPACE;
  }
  TN_REQUIRE(tol >= 0 && max_restarts >= 1, "tn_lanczos_lm1: bad tol/max_restarts");
  const int m = (int)std::min<long long>(std::max(ncv, 2), std::min<long long>(n, kMaxNcv));
  const long long ldv = (n + 1) / 2 * 2;
  Carver cw(workspace, workspace_bytes);
  double* V = cw.take<double>((size_t)ldv * (m + 2));
  double* partial = cw.take<double>((size_t)(m + 1) * dot_chunks(n));
  LanczosState* st = cw.take<LanczosState>(1);
  unsigned* counter = cw.take<unsigned>(1);
  TN_REQUIRE(V && partial && st && counter, "tn_lanczos_lm1: workspace carve failed");
  double* ytmp = V + (size_t)ldv * (m + 1);
  auto vec = [&](int j) { return V + (size_t)ldv * j; };

  TN_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
  TN_CUDA(cudaMemsetAsync(st, 0, sizeof(LanczosState), stream));
  // v_0 = v0 / |v0|
  TN_CUDA(cudaMemcpyAsync(vec(0), v0, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
  TN_CHECK(launch_multidot(vec(0), ldv, 1, vec(0), n, &st->nrm2, partial, counter, stream));
  lanczos_init_kernel<<<1, 1, 0, stream>>>(st);
  TN_LAUNCHED();
  TN_CHECK(launch_scale_dev(vec(0), &st->inv_beta, n, stream));

  int n_matvec = 0;
  int j0 = 0;
  LanczosStatus hs{};
  int status = TN_ERR_NOCONV;
  for (int cycle = 0; cycle < max_restarts; ++cycle) {
    for (int j = j0; j < m; ++j) {
      double* w = vec(j + 1);
      TN_CHECK(tn_effh_matvec(plan, vec(j), w, 0.0, 1.0, stream));
      ++n_matvec;
      if (allreduce) {
        int rc = allreduce(w, n, allreduce_user, stream);
        TN_REQUIRE(rc == 0, "tn_lanczos_lm1: all-reduce callback failed (%d)", rc);
      }
      // CGS2 against v_0..v_j
      TN_CHECK(launch_multidot(V, ldv, j + 1, w, n, st->h, partial, counter, stream));
      TN_CHECK(launch_multi_axpy(w, V, ldv, j + 1, st->h, n, stream));
      TN_CHECK(launch_multidot(V, ldv, j + 1, w, n, st->h2, partial, counter, stream));
      TN_CHECK(launch_multi_axpy(w, V, ldv, j + 1, st->h2, n, stream));
      TN_CHECK(launch_multidot(w, ldv, 1, w, n, &st->nrm2, partial, counter, stream));
      lanczos_finish_step_kernel<<<1, 1, 0, stream>>>(st, j);
      TN_LAUNCHED();
      TN_CHECK(launch_scale_dev(w, &st->inv_beta, n, stream));
    }
    lanczos_ritz_kernel<<<1, 32, 0, stream>>>(st, m, tau, tol);
    TN_LAUNCHED();
    TN_CUDA(cudaMemcpyAsync(&hs, &st->theta, sizeof(LanczosStatus), cudaMemcpyDeviceToHost, stream));
    TN_CUDA(cudaStreamSynchronize(stream));
    if (hs.ql_fail) {
      set_error("tn_lanczos_lm1: tridiagonal QL did not converge");
      return TN_ERR_NOCONV;
    }
    // Ritz vector y = V u
    TN_CHECK(launch_combine(ytmp, V, ldv, hs.m_eff, st->u, n, stream));
    if (hs.converged || m >= n) {
      status = TN_OK;
      break;
    }
    if (cycle + 1 == max_restarts) break;
    // thick restart: v_0 = y, v_1 = residual direction (old v_m), T[0,0] = theta, T[0,1] = s
    TN_CUDA(cudaMemcpyAsync(vec(0), ytmp, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    TN_CUDA(cudaMemcpyAsync(vec(1), vec(m), sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    lanczos_restart_kernel<<<1, 1, 0, stream>>>(st);
    TN_LAUNCHED();
    j0 = 1;
  }
  // normalise and return
  TN_CHECK(launch_multidot(ytmp, ldv, 1, ytmp, n, &st->nrm2, partial, counter, stream));
  lanczos_init_kernel<<<1, 1, 0, stream>>>(st);
  TN_LAUNCHED();
  TN_CHECK(launch_scale_dev(ytmp, &st->inv_beta, n, stream));
  TN_CUDA(cudaMemcpyAsync(vec_out, ytmp, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
  TN_CUDA(cudaStreamSynchronize(stream));
  if (lambda_out) *lambda_out = hs.lambda;
  if (resid_out) *resid_out = hs.resid;
  if (n_matvec_out) *n_matvec_out = n_matvec;
  if (status == TN_ERR_NOCONV) set_error("tn_lanczos_lm1: not converged after %d restart cycles (residual %.3e)", max_restarts, hs.resid);
  return status;
}
