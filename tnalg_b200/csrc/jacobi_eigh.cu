// Symmetric eigenproblem on the one-sided Jacobi kernels (north_star kernel 3 "Jacobi SVD/eigh"): for A = A^T the singular
// triplets (s_i, u_i, v_i) of tn_svd_jacobi are eigenpairs up to a sign, lambda_i = s_i * sign(u_i . v_i).  Used by the
// dense eig_way = 0 local solve (MPSClass.py:792-794: eigs of the explicit matrix 1 - tau*H_eff) and by cross-checks.
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace tn {

// lam[i] = S[i] * sign(sum_r U[r,i] * Vt[i,r])   (one CTA per eigenvalue)
__global__ void eigh_sign_kernel(const double* __restrict__ U, const double* __restrict__ Vt, const double* __restrict__ S, int n,
                                 double* __restrict__ lam) {
  const int i = blockIdx.x;
  double acc = 0.0;
  for (int r = threadIdx.x; r < n; r += blockDim.x) acc += U[(size_t)r * n + i] * Vt[(size_t)i * n + r];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    lam[i] = s >= 0.0 ? S[i] : -S[i];
  }
}

// V[r, j] = U[r, perm[j]]
__global__ void eigh_gather_kernel(const double* __restrict__ U, const int* __restrict__ perm, int n, double* __restrict__ V) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)n * n; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / n), j = (int)(e % n);
    V[e] = U[(size_t)r * n + perm[j]];
  }
}

}  // namespace tn

using namespace tn;

extern "C" size_t tn_eigh_workspace_bytes(int n) {
  return tn_svd_workspace_bytes(n, n) + 2 * align_up(sizeof(double) * (size_t)n * n) + 2 * align_up(sizeof(double) * (size_t)n) +
         align_up(sizeof(int) * (size_t)n) + 1024;
}

// A (n,n) symmetric row-major -> w (n) ascending eigenvalues, V (n,n) row-major with eigenvector j in column j.  Blocking.
extern "C" int tn_eigh_jacobi(const double* A, int n, double* w, double* V, int* sweeps_out, void* workspace, size_t workspace_bytes,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(A && w && V && n > 0, "tn_eigh_jacobi: bad arguments");
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_eigh_jacobi: workspace must be 256-byte aligned");
  if (workspace_bytes < tn_eigh_workspace_bytes(n)) {
    set_error("tn_eigh_jacobi: workspace %zu < %zu bytes", workspace_bytes, tn_eigh_workspace_bytes(n));
    return TN_ERR_WORKSPACE;
  }
  Carver cw(workspace, workspace_bytes);
  double* U = cw.take<double>((size_t)n * n);
  double* Vt = cw.take<double>((size_t)n * n);
  double* S = cw.take<double>(n);
  double* lam = cw.take<double>(n);
  int* perm = cw.take<int>(n);
  TN_REQUIRE(U && Vt && S && lam && perm, "tn_eigh_jacobi: workspace carve failed");
  char* rest = cw.base + cw.used;
  const int st = tn_svd_jacobi(A, n, n, n, U, S, Vt, sweeps_out, rest, workspace_bytes - cw.used, stream);
  if (st != TN_OK && st != TN_ERR_NOCONV) return st;
  eigh_sign_kernel<<<n, 128, 0, stream>>>(U, Vt, S, n, lam);
  TN_LAUNCHED();
  std::vector<double> hl(n);
  TN_CUDA(cudaMemcpyAsync(hl.data(), lam, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
  TN_CUDA(cudaStreamSynchronize(stream));
  std::vector<int> hp(n);
  std::iota(hp.begin(), hp.end(), 0);
  std::stable_sort(hp.begin(), hp.end(), [&](int x, int y) { return hl[x] < hl[y]; });
  std::vector<double> sorted(n);
  for (int i = 0; i < n; ++i) sorted[i] = hl[hp[i]];
  TN_CUDA(cudaMemcpyAsync(perm, hp.data(), sizeof(int) * n, cudaMemcpyHostToDevice, stream));
  TN_CUDA(cudaMemcpyAsync(w, sorted.data(), sizeof(double) * n, cudaMemcpyHostToDevice, stream));
  eigh_gather_kernel<<<(int)std::min<long long>(((long long)n * n + 255) / 256, 4096), 256, 0, stream>>>(U, perm, n, V);
  TN_LAUNCHED();
  TN_CUDA(cudaStreamSynchronize(stream));
  return st;
}
