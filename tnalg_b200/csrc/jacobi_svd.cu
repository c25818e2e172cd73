// One-sided (Hestenes) Jacobi SVD in FP64  (a7 'svd' branch, a11, a12).
//
// The rows of a work matrix W are rotated pairwise until they are mutually orthogonal:
//   m >= n : W = A^T (n rows of length m), the accumulated rotations give V^T, U = normalised rows
//   m <  n : W = A   (m rows of length n), the accumulated rotations give U^T, V^T = normalised rows
// so every dot product and every rotation streams over contiguous memory (coalesced, HBM/L2-bound).
// One kernel launch per round of the round-robin (chess tournament) ordering: r/2 disjoint pairs, one CTA per
// pair, warp-shuffle reductions for (|p|^2, |q|^2, p.q).  The sweep loop stops when a whole sweep applied no
// rotation.  Singular values are sorted on the host (k doubles), the gather of the k_keep leading triplets is a
// kernel.  Jacobi gives high relative accuracy for the small Schmidt values the entanglement spectrum needs.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "common.cuh"
#include "host_math.h"

namespace tn {

constexpr int kJacThreads = 256;

__device__ __forceinline__ double warp_sum_j(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// W: rows x len (row pitch ldw), Acc: rows x rows_acc (row pitch lda).  round in [0, rows-1), rows even.
__global__ void __launch_bounds__(kJacThreads) jacobi_round_kernel(double* __restrict__ W, long long ldw, int len,
                                                                   double* __restrict__ Acc, long long lda, int acc_len,
                                                                   int rows, int round, double tol, unsigned* n_rot) {
  const int i = blockIdx.x;  // pair index
  const int nm1 = rows - 1;
  int p, q;
  if (i == 0) {
    p = round % nm1;
    q = nm1;
  } else {
    p = (round + i) % nm1;
    q = (round + nm1 - i) % nm1;
  }
  if (p > q) { int tmp = p; p = q; q = tmp; }
  double* wp = W + (long long)p * ldw;
  double* wq = W + (long long)q * ldw;
  double a = 0.0, b = 0.0, g = 0.0;
  for (int e = threadIdx.x; e < len; e += kJacThreads) {
    const double x = wp[e], y = wq[e];
    a += x * x;
    b += y * y;
    g += x * y;
  }
  a = warp_sum_j(a); b = warp_sum_j(b); g = warp_sum_j(g);
  __shared__ double red[3][kJacThreads / 32];
  __shared__ double cs[2];
  __shared__ int do_rot;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = g; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double A_ = 0, B_ = 0, G_ = 0;
#pragma unroll
    for (int k = 0; k < kJacThreads / 32; ++k) { A_ += red[0][k]; B_ += red[1][k]; G_ += red[2][k]; }
    int rot = (A_ > 0.0 && B_ > 0.0 && fabs(G_) > tol * sqrt(A_) * sqrt(B_)) ? 1 : 0;
    if (rot) {
      double c, s;
      jacobi_rotation(A_, B_, G_, &c, &s);
      cs[0] = c; cs[1] = s;
      atomicAdd(n_rot, 1u);
    }
    do_rot = rot;
  }
  __syncthreads();
  if (!do_rot) return;
  const double c = cs[0], s = cs[1];
  for (int e = threadIdx.x; e < len; e += kJacThreads) {
    const double x = wp[e], y = wq[e];
    wp[e] = c * x - s * y;
    wq[e] = s * x + c * y;
  }
  if (Acc) {
    double* ap = Acc + (long long)p * lda;
    double* aq = Acc + (long long)q * lda;
    for (int e = threadIdx.x; e < acc_len; e += kJacThreads) {
      const double x = ap[e], y = aq[e];
      ap[e] = c * x - s * y;
      aq[e] = s * x + c * y;
    }
  }
}

// out (rows_out x cols_out, pitch ldo) = in^T, tiled through shared memory
__global__ void transpose_kernel(const double* __restrict__ in, int rows_in, int cols_in, double* __restrict__ out, long long ldo) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int gi = by + r, gj = bx + threadIdx.x;
    if (gi < rows_in && gj < cols_in) tile[r][threadIdx.x] = in[(long long)gi * cols_in + gj];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int gi = bx + r, gj = by + threadIdx.x;  // out[gi][gj] = in[gj][gi]
    if (gi < cols_in && gj < rows_in) out[(long long)gi * ldo + gj] = tile[threadIdx.x][r];
  }
}

__global__ void set_identity_kernel(double* M, int rows, long long ld) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * ld; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / ld, c = i % ld;
    M[i] = (r == c) ? 1.0 : 0.0;
  }
}

__global__ void __launch_bounds__(kJacThreads) row_norms_kernel(const double* __restrict__ W, long long ldw, int len, double* __restrict__ sig) {
  const double* w = W + (long long)blockIdx.x * ldw;
  double a = 0.0;
  for (int e = threadIdx.x; e < len; e += kJacThreads) a += w[e] * w[e];
  a = warp_sum_j(a);
  __shared__ double red[kJacThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < kJacThreads / 32; ++k) s += red[k];
    sig[blockIdx.x] = sqrt(s);
  }
}

// dst row jj (length len, pitch ldd) = scale(jj) * src row perm[jj];  scale = 1/sig[perm[jj]] when normalise != 0
__global__ void gather_rows_kernel(const double* __restrict__ src, long long lds, const int* __restrict__ perm, const double* __restrict__ sig,
                                   int normalise, int len, double* __restrict__ dst, long long ldd) {
  const int jj = blockIdx.x, r = perm[jj];
  double sc = 1.0;
  if (normalise) sc = sig[r] > 1e-300 ? 1.0 / sig[r] : 0.0;
  for (int e = threadIdx.x; e < len; e += blockDim.x) dst[(long long)jj * ldd + e] = sc * src[(long long)r * lds + e];
}

// dst (len x k, row-major) column jj = scale(jj) * src row perm[jj]   (transposing gather)
__global__ void gather_cols_kernel(const double* __restrict__ src, long long lds, const int* __restrict__ perm, const double* __restrict__ sig,
                                   int normalise, int len, int k, double* __restrict__ dst) {
  __shared__ double tile[32][33];
  const int j0 = blockIdx.x * 32, e0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int jj = j0 + r, e = e0 + threadIdx.x;
    double v = 0.0;
    if (jj < k && e < len) {
      const int row = perm[jj];
      double sc = 1.0;
      if (normalise) sc = sig[row] > 1e-300 ? 1.0 / sig[row] : 0.0;
      v = sc * src[(long long)row * lds + e];
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int e = e0 + r, jj = j0 + threadIdx.x;
    if (e < len && jj < k) dst[(long long)e * k + jj] = tile[threadIdx.x][r];
  }
}

__global__ void gather_sig_kernel(const double* sig, const int* perm, int k, double* S) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) S[i] = sig[perm[i]];
}

}  // namespace tn

using namespace tn;

static void svd_dims(int m, int n, int* rows, int* rows_pad, int* len) {
  *rows = std::min(m, n);
  *len = std::max(m, n);
  *rows_pad = (*rows + 1) / 2 * 2;
}

extern "C" size_t tn_svd_workspace_bytes(int m, int n) {
  int rows, rp, len;
  svd_dims(m, n, &rows, &rp, &len);
  size_t ldw = (size_t)(len + 1) / 2 * 2;
  return align_up(sizeof(double) * (size_t)rp * ldw) + align_up(sizeof(double) * (size_t)rp * rp) + align_up(sizeof(double) * rp) +
         align_up(sizeof(int) * (size_t)rp) + align_up(sizeof(unsigned)) + 1024;
}

extern "C" int tn_svd_jacobi(const double* A, int m, int n, int k_keep, double* U, double* S, double* Vt, int* sweeps_out,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(A && S && m > 0 && n > 0, "tn_svd_jacobi: bad arguments");
  int rows, rp, len;
  svd_dims(m, n, &rows, &rp, &len);
  TN_REQUIRE(k_keep >= 1 && k_keep <= rows, "tn_svd_jacobi: k_keep=%d not in 1..%d", k_keep, rows);
  TN_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tn_svd_jacobi: workspace must be 256-byte aligned");
  if (workspace_bytes < tn_svd_workspace_bytes(m, n)) {
    set_error("tn_svd_jacobi: workspace %zu < %zu bytes", workspace_bytes, tn_svd_workspace_bytes(m, n));
    return TN_ERR_WORKSPACE;
  }
  const long long ldw = (long long)(len + 1) / 2 * 2;
  Carver cw(workspace, workspace_bytes);
  double* W = cw.take<double>((size_t)rp * ldw);
  double* Acc = cw.take<double>((size_t)rp * rp);
  double* sig = cw.take<double>(rp);
  int* perm = cw.take<int>(rp);
  unsigned* n_rot = cw.take<unsigned>(1);
  TN_REQUIRE(W && Acc && sig && perm && n_rot, "tn_svd_jacobi: workspace carve failed");
  const bool tall = m >= n;  // W = A^T
  const bool need_acc = tall ? (Vt != nullptr) : (U != nullptr);

  TN_CUDA(cudaMemsetAsync(W, 0, sizeof(double) * (size_t)rp * ldw, stream));
  if (tall) {
    dim3 grid((n + 31) / 32, (m + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, stream>>>(A, m, n, W, ldw);
    TN_LAUNCHED();
  } else {
    TN_CUDA(cudaMemcpy2DAsync(W, sizeof(double) * ldw, A, sizeof(double) * n, sizeof(double) * n, m, cudaMemcpyDeviceToDevice, stream));
  }
  if (need_acc) {
    set_identity_kernel<<<std::min(1024, (rp * rp + 255) / 256), 256, 0, stream>>>(Acc, rp, rp);
    TN_LAUNCHED();
  }
  const double tol = std::max(1e-15, std::sqrt((double)len) * 2.2e-16);
  int sweeps = 0;
  const int max_sweeps = 80;  // QR-preconditioned inputs (ops.CudaBackend.svd) need ~8; raw ill-conditioned ones 20-50
  bool converged = rp < 2;
  while (!converged && sweeps < max_sweeps) {
    TN_CUDA(cudaMemsetAsync(n_rot, 0, sizeof(unsigned), stream));
    for (int round = 0; round < rp - 1; ++round) {
      jacobi_round_kernel<<<rp / 2, kJacThreads, 0, stream>>>(W, ldw, len, need_acc ? Acc : nullptr, rp, rp, rp, round, tol, n_rot);
      TN_LAUNCHED();
    }
    unsigned h_rot = 0;
    TN_CUDA(cudaMemcpyAsync(&h_rot, n_rot, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    TN_CUDA(cudaStreamSynchronize(stream));
    ++sweeps;
    converged = (h_rot == 0);
  }
  if (sweeps_out) *sweeps_out = sweeps;
  row_norms_kernel<<<rp, kJacThreads, 0, stream>>>(W, ldw, len, sig);
  TN_LAUNCHED();
  std::vector<double> hs(rp);
  TN_CUDA(cudaMemcpyAsync(hs.data(), sig, sizeof(double) * rp, cudaMemcpyDeviceToHost, stream));
  TN_CUDA(cudaStreamSynchronize(stream));
  std::vector<int> hp(rows);
  std::iota(hp.begin(), hp.end(), 0);
  std::stable_sort(hp.begin(), hp.end(), [&](int x, int y) { return hs[x] > hs[y]; });
  TN_CUDA(cudaMemcpyAsync(perm, hp.data(), sizeof(int) * rows, cudaMemcpyHostToDevice, stream));
  gather_sig_kernel<<<(k_keep + 255) / 256, 256, 0, stream>>>(sig, perm, k_keep, S);
  TN_LAUNCHED();
  dim3 tb(32, 8);
  if (tall) {
    // U (m x k_keep) columns = normalised rows of W; Vt (k_keep x n) rows = rows of Acc
    if (U) {
      dim3 grid((k_keep + 31) / 32, (m + 31) / 32);
      gather_cols_kernel<<<grid, tb, 0, stream>>>(W, ldw, perm, sig, 1, m, k_keep, U);
      TN_LAUNCHED();
    }
    if (Vt) {
      gather_rows_kernel<<<k_keep, 256, 0, stream>>>(Acc, rp, perm, sig, 0, n, Vt, n);
      TN_LAUNCHED();
    }
  } else {
    // U (m x k_keep) columns = rows of Acc (U = Acc^T); Vt rows = normalised rows of W
    if (U) {
      dim3 grid((k_keep + 31) / 32, (m + 31) / 32);
      gather_cols_kernel<<<grid, tb, 0, stream>>>(Acc, rp, perm, sig, 0, m, k_keep, U);
      TN_LAUNCHED();
    }
    if (Vt) {
      gather_rows_kernel<<<k_keep, 256, 0, stream>>>(W, ldw, perm, sig, 1, n, Vt, n);
      TN_LAUNCHED();
    }
  }
  TN_CUDA(cudaStreamSynchronize(stream));  // hp / hs are host temporaries
  if (!converged) {
    set_error("tn_svd_jacobi: not converged after %d sweeps", sweeps);
    return TN_ERR_NOCONV;
  }
  return TN_OK;
}
