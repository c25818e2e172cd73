// Memory-bound helper kernels shared by the plan, the Lanczos driver and the SVD (HBM roofline class).
#pragma once
#include "common.cuh"

namespace tn {

struct SiteOp {
  double m[kMaxPhys * kMaxPhys];
};

// out[a,s,b] = c_id * x[a,s,b] + c_op * sum_s' op[s,s'] x[a,s',b]     (x is (a,d,b) C-order)
int launch_site_op_axpby(double* out, const double* x, long long a, int d, long long b, double c_id, double c_op,
                         const SiteOp& op, cudaStream_t stream);

// deterministic dot products: result[i] = sum_e V[i*ldv + e] * w[e], i < nvec.  partial: nvec * dot_chunks(n) doubles,
// counter: one unsigned zero-initialised once (the kernel resets it).  Optional fused bookkeeping is done by callers
// in a follow-up single-thread kernel.
int dot_chunks(long long n);
// skip_flag (device int, may be null): when *skip_flag == 0 the kernel returns at once (conditional second CGS pass)
int launch_multidot(const double* V, long long ldv, int nvec, const double* w, long long n, double* result, double* partial,
                    unsigned* counter, cudaStream_t stream, const int* skip_flag = nullptr);
// w[e] -= sum_i h[i] * V[i*ldv + e]
int launch_multi_axpy(double* w, const double* V, long long ldv, int nvec, const double* h, long long n, cudaStream_t stream,
                      const int* skip_flag = nullptr);
// y[e] = sum_i u[i] * V[i*ldv + e]
int launch_combine(double* y, const double* V, long long ldv, int nvec, const double* u, long long n, cudaStream_t stream);
// overflow-safe scaling: scale2[0] = 2^-e, scale2[1] = 2^e with 2^(e-1) <= max|x| < 2^e (both 1 for a zero or non-finite x).
// slot: one zero-initialised unsigned long long of scratch (the kernel leaves it zero again).  Exact, order independent.
int launch_pow2_scale(const double* x, long long n, double* scale2, unsigned long long* slot, cudaStream_t stream);
// x[e] *= *scale (device scalar)
int launch_scale_dev(double* x, const double* scale, long long n, cudaStream_t stream);

}  // namespace tn
