// Symmetric eigenproblem on the one-sided Jacobi kernels (north_star kernel 3 "Jacobi SVD/eigh"): A + sigma*1 with sigma a
// Gershgorin bound is positive semi-definite, so its singular triplets from tn_svd_jacobi ARE its eigenpairs
// (lambda_i = s_i - sigma, eigenvector u_i); no +-|lambda| pairs with ill-defined singular vectors.  Absolute accuracy
// eps * |A|.  Used by the dense eig_way = 0 local solve (MPSClass.py:792-794: eigs of the explicit matrix 1 - tau*H_eff,
// which is positive definite anyway) and by cross-checks.
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace tn {

// rowsum[i] = sum_j |A[i,j]|   (one CTA per row; the Gershgorin radius is the max, taken on the host)
__global__ void eigh_rowsum_kernel(const double* __restrict__ A, int n, double* __restrict__ rowsum) {
  const int i = blockIdx.x;
  double acc = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) acc += fabs(A[(size_t)i * n + c]);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    rowsum[i] = s;
  }
}

// B = (A + A^T)/2 + sigma * 1
__global__ void eigh_shift_kernel(const double* __restrict__ A, int n, double sigma, double* __restrict__ B) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)n * n; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / n), c = (int)(e % n);
    B[e] = 0.5 * (A[e] + A[(size_t)c * n + r]) + (r == c ? sigma : 0.0);
  }
}

// w[j] = S[perm[j]] - sigma
__global__ void eigh_values_kernel(const double* __restrict__ S, const int* __restrict__ perm, int n, double sigma, double* __restrict__ w) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) w[j] = S[perm[j]] - sigma;
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
  const int grid = (int)std::min<long long>(((long long)n * n + 255) / 256, 4096);
  // Gershgorin shift: sigma >= max_i sum_j |A_ij| >= -lambda_min
  eigh_rowsum_kernel<<<n, 128, 0, stream>>>(A, n, lam);
  TN_LAUNCHED();
  std::vector<double> hl(n);
  TN_CUDA(cudaMemcpyAsync(hl.data(), lam, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
  TN_CUDA(cudaStreamSynchronize(stream));
  const double sigma = *std::max_element(hl.begin(), hl.end());
  eigh_shift_kernel<<<grid, 256, 0, stream>>>(A, n, sigma, Vt);  // Vt doubles as the shifted copy (the SVD does not write it: Vt = NULL)
  TN_LAUNCHED();
  const int st = tn_svd_jacobi(Vt, n, n, n, U, S, nullptr, sweeps_out, rest, workspace_bytes - cw.used, stream);
  if (st != TN_OK && st != TN_ERR_NOCONV) return st;
  // S is sorted in decreasing order: ascending eigenvalues = reversed order
  std::vector<int> hp(n);
  for (int i = 0; i < n; ++i) hp[i] = n - 1 - i;
  TN_CUDA(cudaMemcpyAsync(perm, hp.data(), sizeof(int) * n, cudaMemcpyHostToDevice, stream));
  eigh_values_kernel<<<(n + 255) / 256, 256, 0, stream>>>(S, perm, n, sigma, w);
  TN_LAUNCHED();
  eigh_gather_kernel<<<grid, 256, 0, stream>>>(U, perm, n, V);
  TN_LAUNCHED();
  TN_CUDA(cudaStreamSynchronize(stream));
  return st;
}
