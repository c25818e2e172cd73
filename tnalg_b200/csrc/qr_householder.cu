// Blocked Householder QR of the gauge moves (a7: left2right/right2left_decompose_tensor with way='qr',
// TensorBasicModule.py:342-345,378-380 -> np.linalg.qr; every sweep step of DMRG_anyH.py:47-64 calls it).
//
//   A (m, n) = Q (m, k) R (k, n),  k = min(m, n),  Q^T Q = 1 to rounding for ANY conditioning (centre tensors carry the
//   Schmidt spectrum, kappa up to 1e12 and beyond: Gram-matrix / Cholesky variants are not acceptable here).
//   Same reflector convention as LAPACK dgeqr2/dlarfg (beta = -sign(alpha) |x|), so R agrees with np.linalg.qr including
//   the signs of its rows whenever A has full column rank.
//
// Structure (right-looking, panel width 32):
//   qr_panel_kernel   one thread-block CLUSTER per panel.  The panel rows are spread over the CTAs of the cluster and stay in
//                     shared memory for all 32 columns.  Per column ONE cluster-wide reduction (distributed shared memory +
//                     barrier.cluster) yields the whole row  g_k = sum_{r>=j} P[r,j] P[r,k]  of the panel Gram matrix, from
//                     which the column norm (k = j), the reflector products v^T a_k (k > j) and the entries V_k^T v_j of the
//                     compact-WY factor T (k < j) all follow:  y_k = (g_k - beta P[j,k]) / (alpha - beta).
//                     Latency per column = one CTA barrier + one cluster barrier (~0.5 us), not a chain of launches.
//   qr_apply_kernel   C <- (1 - V op(T) V^T) C on a strip of 16 columns per CTA with FP64 tensor-core MMAs (DMMA.8x8x4):
//                     W = V^T C (warps split the rows, fixed-order reduction through shared memory), W <- op(T) W,
//                     C -= V W.  Used for the trailing update (op(T) = T^T) and for forming Q (op(T) = T, panels in reverse).
// No atomics anywhere: results are bit-reproducible (replicas on several GPUs stay identical).
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace tn {

namespace {
constexpr int QNB = 32;        // panel width
constexpr int QTHREADS = 512;  // 16 warps
constexpr int QWARPS = QTHREADS / 32;
constexpr int QMAXC = 8;       // portable cluster size
constexpr int QSLOT = 2 * QNB; // doubles one CTA contributes per column: g[32], row j[32]
constexpr int QMAX_ROWS_PER_CTA = 768;
constexpr int QNC = 16;        // strip width of the apply kernel

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum_q(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

size_t panel_smem_bytes(int R) {
  return sizeof(double) * ((size_t)R * QNB + QWARPS * QNB + 2 * QMAXC * QSLOT + 2 * QNB * QNB + QNB);
}
}  // namespace

// W: work matrix (m x n, leading dimension ld); panel = columns [j0, j0 + nbp), rows [j0, m)
__global__ void __launch_bounds__(QTHREADS, 1) qr_panel_kernel(double* __restrict__ W, int ld, int m, int j0, int nbp,
                                                               double* __restrict__ tau_out, double* __restrict__ T_out, int R) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), c = (int)cluster.block_rank();
  extern __shared__ __align__(16) double sm[];
  double* P = sm;                          // R x 32, this CTA's rows of the panel
  double* red = P + (size_t)R * QNB;       // 16 x 32 cross-warp partial sums
  double* slots = red + QWARPS * QNB;      // [parity][source CTA][64]
  double* Z = slots + 2 * QMAXC * QSLOT;   // Z[k*32 + j] = V_k^T v_j (k < j)
  double* Ts = Z + QNB * QNB;              // compact-WY factor
  double* taus = Ts + QNB * QNB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows_total = m - j0;
  const int row_lo = c * R;
  const int nloc = max(0, min(R, rows_total - row_lo));

  for (int idx = tid; idx < nloc * QNB; idx += QTHREADS) {
    const int r = idx >> 5, k = idx & 31;
    P[idx] = k < nbp ? W[(size_t)(j0 + row_lo + r) * ld + j0 + k] : 0.0;
  }
  for (int idx = tid; idx < 2 * QNB * QNB + QNB; idx += QTHREADS) Z[idx] = 0.0;  // Z, Ts, taus are contiguous
  __syncthreads();
  if (C > 1) cluster.sync();  // every CTA of the cluster is resident before remote shared memory is written

  for (int j = 0; j < nbp; ++j) {
    // ---- row j of the panel Gram matrix over this CTA's rows with panel index >= j ----
    const int r_begin = max(0, j - row_lo);
    {
      double acc = 0.0, acc1 = 0.0;
      int r = r_begin + warp;
      for (; r + QWARPS < nloc; r += 2 * QWARPS) {
        acc += P[r * QNB + j] * P[r * QNB + lane];
        acc1 += P[(r + QWARPS) * QNB + j] * P[(r + QWARPS) * QNB + lane];
      }
      if (r < nloc) acc += P[r * QNB + j] * P[r * QNB + lane];
      red[warp * QNB + lane] = acc + acc1;
    }
    __syncthreads();
    const int par = (j & 1) * QMAXC * QSLOT;
    if (warp == 0) {
      double g = 0.0;
#pragma unroll
      for (int w = 0; w < QWARPS; ++w) g += red[w * QNB + lane];
      const double rj = (c == 0) ? P[j * QNB + lane] : 0.0;  // the diagonal block lives in CTA 0 (R >= 32 when C > 1)
      if (C == 1) {
        slots[par + lane] = g;
        slots[par + QNB + lane] = rj;
      } else {
        for (int dst = 0; dst < C; ++dst) {
          double* remote = cluster.map_shared_rank(slots, dst);
          remote[par + c * QSLOT + lane] = g;
          remote[par + c * QSLOT + QNB + lane] = rj;
        }
      }
    }
    if (C > 1) cluster.sync(); else __syncthreads();
    // ---- every thread: the reduced row (its own column `lane`) and the scalars of column j ----
    double gk = 0.0, gj = 0.0;
    for (int s = 0; s < C; ++s) {
      gk += slots[par + s * QSLOT + lane];
      gj += slots[par + s * QSLOT + j];
    }
    const double rjk = slots[par + QNB + lane], alpha = slots[par + QNB + j];
    double tau, beta, scale;
    if (rows_total - 1 - j == 0 || gj == 0.0) {  // nothing below the diagonal, or a zero column: H = 1 (dlarfg)
      tau = 0.0; beta = alpha; scale = 0.0;
    } else {
      const double nrm = sqrt(gj);
      beta = alpha >= 0.0 ? -nrm : nrm;
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    const double y = (gk - beta * rjk) * scale;  // k > j: v^T a_k;  k < j: V_k^T v_j
    const double ty = tau * y;
    // ---- apply H_j to the columns k > j of this CTA's rows, store v below the diagonal ----
    const int r_start = max(0, j + 1 - row_lo);
    for (int r = r_start + warp; r < nloc; r += QWARPS) {
      const double vr = P[r * QNB + j] * scale;
      __syncwarp();
      if (lane > j) P[r * QNB + lane] -= ty * vr;
      else if (lane == j) P[r * QNB + j] = vr;
    }
    if (c == 0 && warp == (j & (QWARPS - 1))) {  // row j itself (v_j = 1); the warp that owns no other role this column
      if (lane > j) P[j * QNB + lane] -= ty;
      else if (lane == j) P[j * QNB + j] = beta;
      else Z[lane * QNB + j] = y;
      if (lane == j) taus[j] = tau;
    }
    __syncthreads();
  }

  // ---- compact-WY factor: T[j,j] = tau_j, T[0:j, j] = -tau_j T[0:j,0:j] Z[0:j, j]  (dlarft, forward columnwise) ----
  if (c == 0) {
    for (int j = 0; j < nbp; ++j) {
      const double tj = taus[j];
      for (int i = warp; i < j; i += QWARPS) {
        double s = (lane >= i && lane < j) ? Ts[i * QNB + lane] * Z[lane * QNB + j] : 0.0;
        s = warp_sum_q(s);
        if (lane == 0) Ts[i * QNB + j] = -tj * s;
      }
      if (tid == 0) Ts[j * QNB + j] = tj;
      __syncthreads();
    }
    for (int idx = tid; idx < QNB * QNB; idx += QTHREADS) T_out[idx] = Ts[idx];
    if (tid < nbp) tau_out[j0 + tid] = taus[tid];
  }
  for (int idx = tid; idx < nloc * QNB; idx += QTHREADS) {
    const int r = idx >> 5, k = idx & 31;
    if (k < nbp) W[(size_t)(j0 + row_lo + r) * ld + j0 + k] = P[idx];
  }
  if (C > 1) cluster.sync();  // no CTA exits while a peer may still address its shared memory
}

// unit-lower-trapezoidal reflector block stored below the diagonal of the panel columns of Wv
__device__ __forceinline__ double qr_vload(const double* __restrict__ Wv, int ldv, int m, int j0, int nbp, int r, int k) {
  if (r >= m || k >= nbp) return 0.0;
  const int rl = r - j0;
  if (rl < k) return 0.0;
  if (rl == k) return 1.0;
  return Wv[(size_t)r * ldv + j0 + k];
}

// Cm[j0:m, c_begin:c_end] <- (1 - V op(T) V^T) Cm[j0:m, c_begin:c_end];  transT != 0: op(T) = T^T
__global__ void __launch_bounds__(QTHREADS, 1) qr_apply_kernel(const double* __restrict__ Wv, int ldv, int m, int j0, int nbp,
                                                               const double* __restrict__ T, int transT, double* __restrict__ Cm,
                                                               int ldc, int c_begin, int c_end) {
  extern __shared__ __align__(16) double sm_apply[];
  double* Tsm = sm_apply;                                               // 32 x 32
  double (*Wpart)[QNB * QNC] = reinterpret_cast<double (*)[QNB * QNC]>(Tsm + QNB * QNB);  // per-warp partial V^T C
  double* Wfull = Tsm + QNB * QNB + QWARPS * QNB * QNC;
  double* W2 = Wfull + QNB * QNC;                                       // holds -op(T) V^T C
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int c0 = c_begin + blockIdx.x * QNC;
  const int rows_total = m - j0;
  for (int idx = tid; idx < QNB * QNB; idx += QTHREADS) Tsm[idx] = T[idx];

  // ---- phase 1: W = V^T C  (M = 32 reflectors, N = 16 columns, K = rows; each warp owns every 16th group of 4 rows) ----
  double acc[4][QNC / 8][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < QNC / 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  const int n_quads = (rows_total + 3) / 4;
  for (int q = warp; q < n_quads; q += QWARPS) {
    const int r = j0 + 4 * q + t;
    double a[4], b[QNC / 8];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) a[mt] = qr_vload(Wv, ldv, m, j0, nbp, r, 8 * mt + g);
#pragma unroll
    for (int nt = 0; nt < QNC / 8; ++nt) {
      const int cc = c0 + 8 * nt + g;
      b[nt] = (r < m && cc < c_end) ? Cm[(size_t)r * ldc + cc] : 0.0;
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < QNC / 8; ++nt) dmma884(acc[mt][nt], a[mt], b[nt]);
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < QNC / 8; ++nt) {
      Wpart[warp][(8 * mt + g) * QNC + 8 * nt + 2 * t] = acc[mt][nt][0];
      Wpart[warp][(8 * mt + g) * QNC + 8 * nt + 2 * t + 1] = acc[mt][nt][1];
    }
  __syncthreads();
  {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < QWARPS; ++w) s += Wpart[w][tid];  // QNB * QNC == QTHREADS
    Wfull[tid] = s;
  }
  __syncthreads();
  {
    const int i = tid / QNC, cc = tid % QNC;
    double s = 0.0;
    if (transT) {
      for (int k = 0; k <= i; ++k) s += Tsm[k * QNB + i] * Wfull[k * QNC + cc];
    } else {
      for (int k = i; k < QNB; ++k) s += Tsm[i * QNB + k] * Wfull[k * QNC + cc];
    }
    W2[tid] = -s;
  }
  __syncthreads();
  // ---- phase 2: C += V W2  (M = rows, N = 16, K = 32; each warp owns every 16th group of 8 rows) ----
  const int n_oct = (rows_total + 7) / 8;
  for (int o = warp; o < n_oct; o += QWARPS) {
    const int r = j0 + 8 * o + g;
    double cacc[QNC / 8][2];
#pragma unroll
    for (int nt = 0; nt < QNC / 8; ++nt) {
      const int cc = c0 + 8 * nt + 2 * t;
      cacc[nt][0] = (r < m && cc < c_end) ? Cm[(size_t)r * ldc + cc] : 0.0;
      cacc[nt][1] = (r < m && cc + 1 < c_end) ? Cm[(size_t)r * ldc + cc + 1] : 0.0;
    }
#pragma unroll
    for (int ks = 0; ks < QNB / 4; ++ks) {
      const double a = qr_vload(Wv, ldv, m, j0, nbp, r, 4 * ks + t);
#pragma unroll
      for (int nt = 0; nt < QNC / 8; ++nt) dmma884(cacc[nt], a, W2[(4 * ks + t) * QNC + 8 * nt + g]);
    }
#pragma unroll
    for (int nt = 0; nt < QNC / 8; ++nt) {
      const int cc = c0 + 8 * nt + 2 * t;
      if (r < m && cc < c_end) Cm[(size_t)r * ldc + cc] = cacc[nt][0];
      if (r < m && cc + 1 < c_end) Cm[(size_t)r * ldc + cc + 1] = cacc[nt][1];
    }
  }
}

// out (rows x cols, ld ldo) = in^T or in;  tiled through shared memory
__global__ void qr_copy_kernel(const double* __restrict__ in, int ldi, double* __restrict__ out, int ldo, int rows_out, int cols_out,
                               int transpose) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;  // bx: output column block, by: output row block
  if (!transpose) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int r = by + i, c = bx + threadIdx.x;
      if (r < rows_out && c < cols_out) out[(size_t)r * ldo + c] = in[(size_t)r * ldi + c];
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
    if (r < rows_out && c < cols_out) out[(size_t)r * ldo + c] = tile[threadIdx.x][i];
  }
}

// R (k x n) = upper trapezoid of the factored work matrix; Q (m x k) = [1; 0]
__global__ void qr_extract_r_kernel(const double* __restrict__ W, int ld, double* __restrict__ Rm, int k, int n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)k * n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / n), c = (int)(i % n);
    Rm[i] = c >= r ? W[(size_t)r * ld + c] : 0.0;
  }
}
__global__ void qr_eye_kernel(double* __restrict__ Q, int m, int k) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)m * k; i += (long long)gridDim.x * blockDim.x)
    Q[i] = (i / k == i % k) ? 1.0 : 0.0;
}

static int grid_for(long long n) { return (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8); }

static int launch_panel(double* W, int ld, int m, int j0, int nbp, double* tau, double* T, cudaStream_t stream) {
  const int rows = m - j0;
  int C = 1;
  while (C < QMAXC && (rows + C - 1) / C > 256 && rows / (2 * C) >= QNB) C *= 2;
  int R = ((rows + C - 1) / C + 7) / 8 * 8;
  TN_REQUIRE(R <= QMAX_ROWS_PER_CTA, "tn_qr: %d rows exceed the panel capacity (%d rows)", rows, QMAXC * QMAX_ROWS_PER_CTA);
  TN_REQUIRE(C == 1 || R >= QNB, "tn_qr: internal panel split");
  const size_t smem = panel_smem_bytes(R);
  static size_t configured = 0;
  if (smem > configured) {
    TN_CUDA(cudaFuncSetAttribute(qr_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem_bytes(QMAX_ROWS_PER_CTA)));
    configured = panel_smem_bytes(QMAX_ROWS_PER_CTA);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C);
  cfg.blockDim = dim3(QTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TN_CUDA(cudaLaunchKernelEx(&cfg, qr_panel_kernel, W, ld, m, j0, nbp, tau, T, R));
  TN_LAUNCHED();
  return TN_OK;
}

static int launch_apply(const double* Wv, int ldv, int m, int j0, int nbp, const double* T, int transT, double* Cm, int ldc,
                        int c_begin, int c_end, cudaStream_t stream) {
  if (c_end <= c_begin || m - j0 <= 0) return TN_OK;
  const int strips = (c_end - c_begin + QNC - 1) / QNC;
  constexpr size_t smem = sizeof(double) * (QNB * QNB + QWARPS * QNB * QNC + 2 * QNB * QNC);
  static bool configured = false;
  if (!configured) {
    TN_CUDA(cudaFuncSetAttribute(qr_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  qr_apply_kernel<<<strips, QTHREADS, smem, stream>>>(Wv, ldv, m, j0, nbp, T, transT, Cm, ldc, c_begin, c_end);
  TN_LAUNCHED();
  return TN_OK;
}

}  // namespace tn

using namespace tn;

extern "C" size_t tn_qr_workspace_bytes(int m, int n) {
  const int k = std::min(m, n);
  const size_t panels = (size_t)(k + QNB - 1) / QNB;
  return align_up(sizeof(double) * (size_t)m * n) + align_up(sizeof(double) * (size_t)m * k) + align_up(sizeof(double) * panels * QNB * QNB) +
         align_up(sizeof(double) * (size_t)(k + QNB)) + 1024;
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
  TN_REQUIRE(W && Qw && Tall && tau, "tn_qr_householder: workspace carve failed");
  const dim3 tb(32, 8);
  // work copy of A (row-major m x n)
  qr_copy_kernel<<<dim3((n + 31) / 32, (m + 31) / 32), tb, 0, stream>>>(A, trans_in ? m : n, W, n, m, n, trans_in ? 1 : 0);
  TN_LAUNCHED();
  for (int p = 0; p < panels; ++p) {
    const int j0 = p * QNB, nbp = std::min(QNB, k - j0);
    TN_CHECK(launch_panel(W, n, m, j0, nbp, tau, Tall + (size_t)p * QNB * QNB, stream));
    TN_CHECK(launch_apply(W, n, m, j0, nbp, Tall + (size_t)p * QNB * QNB, 1, W, n, j0 + nbp, n, stream));
  }
  qr_extract_r_kernel<<<grid_for((long long)k * n), 256, 0, stream>>>(W, n, R, k, n);
  TN_LAUNCHED();
  double* Qdst = trans_q ? Qw : Q;
  qr_eye_kernel<<<grid_for((long long)m * k), 256, 0, stream>>>(Qdst, m, k);
  TN_LAUNCHED();
  for (int p = panels - 1; p >= 0; --p) {
    const int j0 = p * QNB, nbp = std::min(QNB, k - j0);
    TN_CHECK(launch_apply(W, n, m, j0, nbp, Tall + (size_t)p * QNB * QNB, 0, Qdst, k, j0, k, stream));
  }
  if (trans_q) {
    qr_copy_kernel<<<dim3((m + 31) / 32, (k + 31) / 32), tb, 0, stream>>>(Qw, k, Q, m, k, m, 1);
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
