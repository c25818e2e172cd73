// Small dense-math routines shared by device kernels and CPU unit tests (tests/test_host_math.py builds this
// header with g++).  No CUDA-specific constructs besides the TN_HD qualifier.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TN_HD __host__ __device__
#else
#define TN_HD
#endif

namespace tn {

constexpr int kMaxNcv = 64;

TN_HD inline double tn_sign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// Symmetric tridiagonal eigenproblem by implicit QL with Wilkinson shifts.
//   d[0..n)  diagonal in, eigenvalues out (unsorted);  e[i] couples i and i+1 (e[n-1] ignored, destroyed);
//   z (n x n, row-major, leading dimension ldz) must hold the identity on entry; column k is the eigenvector of d[k].
// Returns 0 on success, 1 when an eigenvalue needed more than 60 iterations.
// Rows k = k0, k0 + kstride, ... of z are updated by the caller: the scalar recurrences are deterministic, so several
// threads may run the routine redundantly on private copies of d/e while each owns a subset of the rows of a shared z
// (lanczos_ritz_kernel: one lane per row).  (k0, kstride) = (0, 1) is the plain serial routine.
TN_HD inline int tridiag_ql_rows(int n, double* d, double* e, double* z, int ldz, int k0, int kstride);
TN_HD inline int tridiag_ql(int n, double* d, double* e, double* z, int ldz) { return tridiag_ql_rows(n, d, e, z, ldz, 0, 1); }

TN_HD inline int tridiag_ql_rows(int n, double* d, double* e, double* z, int ldz, int k0, int kstride) {
  if (n <= 0) return 0;
  e[n - 1] = 0.0;
  for (int l = 0; l < n; ++l) {
    int iter = 0;
    int m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 60) return 1;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + tn_sign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double b = c * e[i];
          // Givens rotation (r, s, c) of (f, g).  hypot + two divisions is a chain of ~40 dependent FP64 operations (FP64
          // results return after ~38 cycles on B200; the QL of a 20 x 20 Lanczos matrix took 230 us per restart cycle): when
          // f^2 + g^2 is safely inside the exponent range one reciprocal square root with a Newton step does it
          const double h2 = f * f + g * g;
          if (h2 > 1e-280 && h2 < 1e280) {
#ifdef __CUDA_ARCH__
            double rn = rsqrt(h2);
#else
            double rn = 1.0 / sqrt(h2);
#endif
            rn = rn * (1.5 - 0.5 * h2 * rn * rn);
            r = h2 * rn;
            e[i + 1] = r;
            s = f * rn;
            c = g * rn;
          } else {
            r = hypot(f, g);
            e[i + 1] = r;
            if (r == 0.0) {
              d[i + 1] -= p;
              e[m] = 0.0;
              break;
            }
            s = f / r;
            c = g / r;
          }
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          for (int k = k0; k < n; k += kstride) {
            f = z[k * ldz + i + 1];
            z[k * ldz + i + 1] = s * z[k * ldz + i] + c * f;
            z[k * ldz + i] = c * z[k * ldz + i] - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return 0;
}

// One-sided Jacobi rotation for the column pair with squared norms (alpha, beta) and inner product gamma:
// returns (c, s) such that  p' = c p - s q,  q' = s p + c q  are orthogonal.
TN_HD inline void jacobi_rotation(double alpha, double beta, double gamma, double* c, double* s) {
  // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)),  zeta = (beta - alpha) / (2 gamma), written with one sqrt, one division
  // and one reciprocal sqrt (the rotation is on the latency-critical path of every Jacobi round)
  const double dd = beta - alpha;
  const double r = sqrt(dd * dd + 4.0 * gamma * gamma);
  double t;
  if (r > 0.0 && r < 1e300) {
    t = 2.0 * gamma / (dd + tn_sign(r, dd));
  } else {  // squares under- or overflowed: the scale-free form
    const double zeta = dd / (2.0 * gamma);
    t = tn_sign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  }
#ifdef __CUDA_ARCH__
  *c = rsqrt(1.0 + t * t);
#else
  *c = 1.0 / sqrt(1.0 + t * t);
#endif
  *s = *c * t;
}

}  // namespace tn
