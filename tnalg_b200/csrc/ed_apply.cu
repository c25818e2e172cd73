// Exact-diagonalisation cross-checker on the device: the full-space operator of
// library/EDspinClass.py:69-77 (EDbasic.project_all_hamilt:  v - tau * sum_n h_{c(n)}(p1_n, p2_n) v  on a d^L vector,
// site 0 the slowest index) as one gather kernel, and the ground state through the same device-resident Lanczos that
// solves the DMRG local problems (exact_ground_state, algorithms/ExactDiagonalizationAlgo.py:12-24, uses eigsh the same way).
// HBM/L2-bound integer-index work: one thread per output element, d^2 reads per coupling.
#include <vector>

#include "common.cuh"

namespace tn {

struct EdTerm {
  long long stride1, stride2;  // d^(L-1-p1), d^(L-1-p2)
  int h;                       // index of the d^2 x d^2 two-site matrix
  int pad;
};

__global__ void ed_apply_kernel(double* __restrict__ out, const double* __restrict__ v, long long n, int d, const EdTerm* __restrict__ terms,
                                int n_terms, const double* __restrict__ hs, double c_id, double c_h) {
  const int dd = d * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int tI = 0; tI < n_terms; ++tI) {
      const EdTerm tm = terms[tI];
      const int s1 = (int)((i / tm.stride1) % d), s2 = (int)((i / tm.stride2) % d);
      const long long base = i - s1 * tm.stride1 - s2 * tm.stride2;
      const double* hrow = hs + (size_t)tm.h * dd * dd + (size_t)(s1 * d + s2) * dd;
      for (int q1 = 0; q1 < d; ++q1)
        for (int q2 = 0; q2 < d; ++q2) acc += hrow[q1 * d + q2] * v[base + q1 * tm.stride1 + q2 * tm.stride2];
    }
    out[i] = c_id * v[i] + c_h * acc;
  }
}

struct EdOp {
  long long n;
  int d, n_terms;
  const EdTerm* terms;
  const double* hs;
};

static int ed_matvec_cb(const double* x, double* y, void* user, void* stream) {
  const EdOp* op = static_cast<const EdOp*>(user);
  const int grid = (int)std::min<long long>((op->n + 255) / 256, (long long)sm_count() * 16);
  ed_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, x, op->n, op->d, op->terms, op->n_terms, op->hs, 0.0, 1.0);
  ++g_launches;
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace tn

using namespace tn;

static int ed_setup(int L, int d, int n_terms, const int* p1, const int* p2, const int* h_index, int n_h, std::vector<EdTerm>* out,
                    long long* n_out) {
  TN_REQUIRE(L >= 2 && d >= 2 && n_terms > 0 && p1 && p2 && h_index && n_h > 0, "tn_ed: bad arguments");
  long long n = 1;
  for (int i = 0; i < L; ++i) {
    n *= d;
    TN_REQUIRE(n <= (1LL << 31), "tn_ed: d^L too large");
  }
  auto stride = [&](int p) {
    long long s = 1;
    for (int i = 0; i < L - 1 - p; ++i) s *= d;
    return s;
  };
  out->resize(n_terms);
  for (int tI = 0; tI < n_terms; ++tI) {
    TN_REQUIRE(p1[tI] >= 0 && p1[tI] < L && p2[tI] >= 0 && p2[tI] < L && p1[tI] != p2[tI], "tn_ed: coupling %d has bad sites", tI);
    TN_REQUIRE(h_index[tI] >= 0 && h_index[tI] < n_h, "tn_ed: coupling %d has a bad Hamiltonian index", tI);
    (*out)[tI] = EdTerm{stride(p1[tI]), stride(p2[tI]), h_index[tI], 0};
  }
  *n_out = n;
  return TN_OK;
}

extern "C" size_t tn_ed_workspace_bytes(int L, int d, int n_terms, int n_h, int ncv) {
  long long n = 1;
  for (int i = 0; i < L; ++i) n *= d;
  return align_up(sizeof(EdTerm) * (size_t)n_terms) + align_up(sizeof(double) * (size_t)n_h * d * d * d * d) +
         tn_lanczos_workspace_bytes(n, ncv) + 1024;
}

// out = c_id * v + c_h * sum_n h[h_index[n]] acting on sites (p1[n], p2[n]);  hs: n_h matrices (d^2 x d^2) row-major, the row /
// column index is (s_p1, s_p2)  -- (c_id, c_h) = (1, -tau) is EDbasic.project_all_hamilt
extern "C" int tn_ed_apply(double* out, const double* v, int L, int d, int n_terms, const int* p1, const int* p2, const int* h_index,
                           const double* hs /* [host] */, int n_h, double c_id, double c_h, void* workspace, size_t workspace_bytes,
                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(out && v && out != v && hs && workspace, "tn_ed_apply: null argument");
  std::vector<EdTerm> terms;
  long long n = 0;
  TN_CHECK(ed_setup(L, d, n_terms, p1, p2, h_index, n_h, &terms, &n));
  Carver cw(workspace, workspace_bytes);
  EdTerm* td = cw.take<EdTerm>(n_terms);
  double* hd = cw.take<double>((size_t)n_h * d * d * d * d);
  TN_REQUIRE(td && hd, "tn_ed_apply: workspace too small");
  TN_CUDA(cudaMemcpyAsync(td, terms.data(), sizeof(EdTerm) * n_terms, cudaMemcpyHostToDevice, stream));
  TN_CUDA(cudaMemcpyAsync(hd, hs, sizeof(double) * (size_t)n_h * d * d * d * d, cudaMemcpyHostToDevice, stream));
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 16);
  ed_apply_kernel<<<grid, 256, 0, stream>>>(out, v, n, d, td, n_terms, hd, c_id, c_h);
  TN_LAUNCHED();
  return TN_OK;
}

// dominant eigenpair of 1 - tau*H on the full d^L space (what eigsh(heff, k=1, which='LM') returns in exact_ground_state)
extern "C" int tn_ed_ground_state(int L, int d, int n_terms, const int* p1, const int* p2, const int* h_index, const double* hs, int n_h,
                                  double tau, const double* v0, double tol, int ncv, int max_restarts, double* lambda_out,
                                  double* vec_out, int* n_matvec_out, double* resid_out, void* workspace, size_t workspace_bytes,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(v0 && vec_out && hs && workspace, "tn_ed_ground_state: null argument");
  TN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_ed_ground_state: workspace must be 256-byte aligned");
  if (workspace_bytes < tn_ed_workspace_bytes(L, d, n_terms, n_h, ncv)) {
    set_error("tn_ed_ground_state: workspace too small");
    return TN_ERR_WORKSPACE;
  }
  std::vector<EdTerm> terms;
  long long n = 0;
  TN_CHECK(ed_setup(L, d, n_terms, p1, p2, h_index, n_h, &terms, &n));
  Carver cw(workspace, workspace_bytes);
  EdTerm* td = cw.take<EdTerm>(n_terms);
  double* hd = cw.take<double>((size_t)n_h * d * d * d * d);
  TN_REQUIRE(td && hd, "tn_ed_ground_state: workspace carve failed");
  TN_CUDA(cudaMemcpyAsync(td, terms.data(), sizeof(EdTerm) * n_terms, cudaMemcpyHostToDevice, stream));
  TN_CUDA(cudaMemcpyAsync(hd, hs, sizeof(double) * (size_t)n_h * d * d * d * d, cudaMemcpyHostToDevice, stream));
  EdOp op{n, d, n_terms, td, hd};
  char* rest = cw.base + cw.used;
  return tn_lanczos_generic(ed_matvec_cb, &op, n, tau, v0, tol, ncv, max_restarts, nullptr, 0, 0, lambda_out, vec_out, n_matvec_out,
                            resid_out, rest, workspace_bytes - cw.used, stream);
}
