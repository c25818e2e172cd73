// Memory-bound kernels: site-operator axpby (the '0_s_0' group and the identity part of the matvec), linear
// combinations of environment matrices, deterministic multi-dot / multi-axpy for Lanczos, trace.
// All are grid-stride, 16-byte vectorised where alignment allows, and judged against the HBM roofline.
#include <vector>

#include "vector_ops.cuh"

namespace tn {

constexpr int kDotChunk = 8192;  // elements per CTA in the dot kernels
constexpr int kDotThreads = 256;

__global__ void site_op_axpby_kernel(double* __restrict__ out, const double* __restrict__ x, long long a, int d, long long b,
                                     double c_id, double c_op, SiteOp op) {
  const long long n = a * d * b;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long ia = i / (d * b);
    const long long rem = i - ia * d * b;
    const int s = (int)(rem / b);
    const long long ib = rem - (long long)s * b;
    double v = c_id * x[i];
    if (c_op != 0.0) {
      double acc = 0.0;
      for (int sp = 0; sp < d; ++sp) acc += op.m[s * d + sp] * x[(ia * d + sp) * b + ib];
      v += c_op * acc;
    }
    out[i] = v;
  }
}

int launch_site_op_axpby(double* out, const double* x, long long a, int d, long long b, double c_id, double c_op,
                         const SiteOp& op, cudaStream_t stream) {
  const long long n = a * d * b;
  int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  site_op_axpby_kernel<<<grid, 256, 0, stream>>>(out, x, a, d, b, c_id, c_op, op);
  TN_LAUNCHED();
  return TN_OK;
}

constexpr int kMaxLincomb = 16;
struct LincombArgs {
  const double* x[kMaxLincomb];
  double c[kMaxLincomb];
  int n_terms;
  int accumulate;
};

__global__ void lincomb_kernel(double* __restrict__ out, long long n, LincombArgs args) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = args.accumulate ? out[i] : 0.0;
#pragma unroll 4
    for (int t = 0; t < args.n_terms; ++t) v += args.c[t] * args.x[t][i];
    out[i] = v;
  }
}

int dot_chunks(long long n) { return (int)((n + kDotChunk - 1) / kDotChunk); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid (chunks, nvec).  Each CTA reduces one chunk of one vector pair with warp shuffles; the last CTA to finish
// (atomic ticket) sums the per-chunk partials in index order, so the result is bit-reproducible.
__global__ void __launch_bounds__(kDotThreads) multidot_kernel(const double* __restrict__ V, long long ldv, const double* __restrict__ w,
                                                               long long n, double* __restrict__ result,
                                                               double* __restrict__ partial, unsigned* counter, const int* skip_flag) {
  if (skip_flag && *skip_flag == 0) return;  // uniform over the grid
  const int chunk = blockIdx.x, vec = blockIdx.y, chunks = gridDim.x;
  const double* v = V + (long long)vec * ldv;
  const long long e0 = (long long)chunk * kDotChunk;
  const long long e1 = min(e0 + kDotChunk, n);
  double acc = 0.0;
  if ((((uintptr_t)(v + e0) | (uintptr_t)(w + e0)) & 15) == 0) {
    // 16-byte loads, four independent pairs in flight per thread (HBM-bound: bytes in flight are what matters)
    const double2* v2 = reinterpret_cast<const double2*>(v + e0);
    const double2* w2 = reinterpret_cast<const double2*>(w + e0);
    const int n2 = (int)((e1 - e0) >> 1);
    double acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int e = threadIdx.x;
    for (; e + 3 * kDotThreads < n2; e += 4 * kDotThreads) {
      const double2 a0 = v2[e], a1 = v2[e + kDotThreads], a2 = v2[e + 2 * kDotThreads], a3 = v2[e + 3 * kDotThreads];
      const double2 b0 = w2[e], b1 = w2[e + kDotThreads], b2 = w2[e + 2 * kDotThreads], b3 = w2[e + 3 * kDotThreads];
      acc += a0.x * b0.x + a0.y * b0.y;
      acc1 += a1.x * b1.x + a1.y * b1.y;
      acc2 += a2.x * b2.x + a2.y * b2.y;
      acc3 += a3.x * b3.x + a3.y * b3.y;
    }
    for (; e < n2; e += kDotThreads) {
      const double2 a0 = v2[e], b0 = w2[e];
      acc += a0.x * b0.x + a0.y * b0.y;
    }
    acc = (acc + acc1) + (acc2 + acc3);
    if (((e1 - e0) & 1) && threadIdx.x == 0) acc += v[e1 - 1] * w[e1 - 1];
  } else {
    for (long long e = e0 + threadIdx.x; e < e1; e += kDotThreads) acc += v[e] * w[e];
  }
  acc = warp_sum(acc);
  __shared__ double s_part[kDotThreads / 32];
  __shared__ bool s_last;
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kDotThreads / 32; ++i) s += s_part[i];
    partial[(long long)vec * chunks + chunk] = s;
    __threadfence();
    unsigned ticket = atomicAdd(counter, 1u);
    s_last = (ticket == (unsigned)(gridDim.x * gridDim.y) - 1u);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    // one warp per vector: lanes take the chunk partials strided (loads in flight instead of a serial chain of 256
    // dependent L2 reads), then a fixed shuffle tree -- the summation order depends only on (chunks), so the result stays
    // bit-reproducible
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < (int)gridDim.y; i += kDotThreads / 32) {
      const volatile double* pp = partial + (long long)i * chunks;
      double s = 0.0;
      for (int c = lane; c < chunks; c += 32) s += pp[c];
      s = warp_sum(s);
      if (lane == 0) result[i] = s;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

int launch_multidot(const double* V, long long ldv, int nvec, const double* w, long long n, double* result, double* partial,
                    unsigned* counter, cudaStream_t stream, const int* skip_flag) {
  dim3 grid(dot_chunks(n), nvec);
  multidot_kernel<<<grid, kDotThreads, 0, stream>>>(V, ldv, w, n, result, partial, counter, skip_flag);
  TN_LAUNCHED();
  return TN_OK;
}

__global__ void multi_axpy_kernel(double* __restrict__ w, const double* __restrict__ V, long long ldv, int nvec,
                                  const double* __restrict__ h, long long n, const int* skip_flag) {
  if (skip_flag && *skip_flag == 0) return;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    double v = w[e];
    for (int i = 0; i < nvec; ++i) v -= h[i] * V[(long long)i * ldv + e];
    w[e] = v;
  }
}

int launch_multi_axpy(double* w, const double* V, long long ldv, int nvec, const double* h, long long n, cudaStream_t stream,
                      const int* skip_flag) {
  int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  multi_axpy_kernel<<<grid, 256, 0, stream>>>(w, V, ldv, nvec, h, n, skip_flag);
  TN_LAUNCHED();
  return TN_OK;
}

__global__ void combine_kernel(double* __restrict__ y, const double* __restrict__ V, long long ldv, int nvec,
                               const double* __restrict__ u, long long n) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int i = 0; i < nvec; ++i) v += u[i] * V[(long long)i * ldv + e];
    y[e] = v;
  }
}

int launch_combine(double* y, const double* V, long long ldv, int nvec, const double* u, long long n, cudaStream_t stream) {
  int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  combine_kernel<<<grid, 256, 0, stream>>>(y, V, ldv, nvec, u, n);
  TN_LAUNCHED();
  return TN_OK;
}

__global__ void scale_dev_kernel(double* __restrict__ x, const double* __restrict__ scale, long long n) {
  const double s = *scale;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) x[e] *= s;
}

int launch_scale_dev(double* x, const double* scale, long long n, cudaStream_t stream) {
  int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  scale_dev_kernel<<<grid, 256, 0, stream>>>(x, scale, n);
  TN_LAUNCHED();
  return TN_OK;
}

__global__ void maxabs_kernel(const double* __restrict__ x, long long n, unsigned long long* slot) {
  double m = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) m = fmax(m, fabs(x[e]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  // non-negative doubles order like their bit patterns (NaN never wins fmax; Inf is the largest pattern)
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(slot, (unsigned long long)__double_as_longlong(m));
}
__global__ void pow2_scale_kernel(unsigned long long* slot, double* scale2) {
  const double m = __longlong_as_double((long long)*slot);
  *slot = 0ull;
  int e = 0;
  if (m > 0.0 && m < 1.7e308) frexp(m, &e);
  scale2[0] = ldexp(1.0, -e);
  scale2[1] = ldexp(1.0, e);
}

int launch_pow2_scale(const double* x, long long n, double* scale2, unsigned long long* slot, cudaStream_t stream) {
  int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  maxabs_kernel<<<grid, 256, 0, stream>>>(x, n, slot);
  TN_LAUNCHED();
  pow2_scale_kernel<<<1, 1, 0, stream>>>(slot, scale2);
  TN_LAUNCHED();
  return TN_OK;
}

__global__ void trace_kernel(const double* __restrict__ E, int n, double* result) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += E[(long long)i * n + i];
  acc = warp_sum(acc);
  __shared__ double s_part[32];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += s_part[i];
    *result = s;
  }
}

}  // namespace tn

using namespace tn;

extern "C" int tn_lincomb(double* out, long long n, int n_terms, const double* const* xs, const double* coeffs, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(out && n > 0 && n_terms > 0 && xs && coeffs, "tn_lincomb: bad arguments");
  int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  for (int t0 = 0; t0 < n_terms; t0 += kMaxLincomb) {
    LincombArgs a;
    a.n_terms = std::min(kMaxLincomb, n_terms - t0);
    a.accumulate = t0 > 0;
    for (int t = 0; t < a.n_terms; ++t) {
      TN_REQUIRE(xs[t0 + t], "tn_lincomb: null term %d", t0 + t);
      a.x[t] = xs[t0 + t];
      a.c[t] = coeffs[t0 + t];
    }
    lincomb_kernel<<<grid, 256, 0, stream>>>(out, n, a);
    TN_LAUNCHED();
  }
  return TN_OK;
}

extern "C" int tn_apply_site_op(double* out, const double* T, int a, int d, int b, const double* op, void* stream_) {
  TN_REQUIRE(out && T && op && a > 0 && b > 0 && d >= 1 && d <= kMaxPhys, "tn_apply_site_op: bad arguments");
  TN_REQUIRE(out != T, "tn_apply_site_op: in-place not supported");
  SiteOp o{};
  for (int i = 0; i < d * d; ++i) o.m[i] = op[i];
  return launch_site_op_axpby(out, T, a, d, b, 0.0, 1.0, o, static_cast<cudaStream_t>(stream_));
}

extern "C" size_t tn_dot_workspace_bytes(long long n) { return align_up(sizeof(double) * (size_t)dot_chunks(n)) + 256; }

extern "C" int tn_dot(const double* x, const double* y, long long n, double* result, void* workspace, size_t workspace_bytes,
                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(x && y && result && n > 0 && workspace, "tn_dot: bad arguments");
  if (workspace_bytes < tn_dot_workspace_bytes(n)) {
    set_error("tn_dot: workspace too small");
    return TN_ERR_WORKSPACE;
  }
  Carver cw(workspace, workspace_bytes);
  unsigned* counter = cw.take<unsigned>(1);
  double* partial = cw.take<double>(dot_chunks(n));
  TN_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
  return launch_multidot(x, n, 1, y, n, result, partial, counter, stream);
}

extern "C" int tn_trace(const double* E, int n, double* result, void* stream_) {
  TN_REQUIRE(E && result && n > 0, "tn_trace: bad arguments");
  trace_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>(E, n, result);
  TN_LAUNCHED();
  return TN_OK;
}
