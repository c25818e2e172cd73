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

// ---------------------------------------------------------------------------------------------------------------------
// Extreme eigenpair of a symmetric tridiagonal without the full QL (the Lanczos Ritz kernel needs ONE pair per restart cycle;
// the QL above is a chain of ~500 dependent Givens rotations, 160 us for a 20 x 20 matrix at FP64 latencies).
//   * eigenvalue: multisection on Sturm counts evaluated with the division-free minor recurrence (one FMA per row on the
//     critical path), many shifts in parallel (one per lane);
//   * eigenvector: twisted factorisation (forward and backward pivot sequences, joined where |gamma_k| is smallest), as in
//     LAPACK's dlar1v.
// The matrix is expected normalised to the Gershgorin disc [-1, 1] (tridiag_normalise), so no minor can overflow for n <= 64.
// ---------------------------------------------------------------------------------------------------------------------

// centre / radius of the Gershgorin interval (radius >= tiny so that a zero matrix stays finite)
TN_HD inline void tridiag_gershgorin(int n, const double* d, const double* e, double* centre, double* radius) {
  double lo = d[0], hi = d[0];
  for (int i = 0; i < n; ++i) {
    const double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i + 1 < n ? fabs(e[i]) : 0.0);
    lo = fmin(lo, d[i] - r);
    hi = fmax(hi, d[i] + r);
  }
  *centre = 0.5 * (lo + hi);
  *radius = fmax(0.5 * (hi - lo), 1e-300);
}

// number of eigenvalues strictly below x: sign changes of the leading principal minors p_0 = 1, p_1 = d_0 - x,
// p_{i+1} = (d_i - x) p_i - e_{i-1}^2 p_{i-1}; a zero minor takes the sign opposite to its predecessor.  e2[i] = e[i]^2.
TN_HD inline int tridiag_count_below(int n, const double* d, const double* e2, double x) {
  double pm = 1.0, p = d[0] - x;
  int neg = (p < 0.0 || p == 0.0) ? 1 : 0;
  int count = neg;
  for (int i = 1; i < n; ++i) {
    double pn = (d[i] - x) * p - e2[i - 1] * pm;
    pm = p;
    p = pn;
    const double a = fabs(p);
    if (a > 1e150 || (a < 1e-150 && fabs(pm) < 1e-150)) {  // keep the pair inside the exponent range (signs are unaffected)
      const double f = a > 1e150 ? 0x1p-600 : 0x1p600;
      p *= f;
      pm *= f;
    }
    const int neg_new = (p < 0.0) ? 1 : (p > 0.0 ? 0 : 1 - neg);
    count += neg_new != neg;
    neg = neg_new;
  }
  return count;
}

// interior point idx (0 .. P-1) of the P-way multisection grid on (lo, hi)
TN_HD inline double multisect_point(double lo, double hi, int P, int idx) { return lo + (hi - lo) * ((idx + 1.0) / (P + 1.0)); }
// `first` = smallest idx whose point has count_below >= k + 1 (P when none has): the bracket count(lo) <= k < count(hi) shrinks
TN_HD inline void multisect_shrink(double* lo, double* hi, int P, int first) {
  const double l = *lo, h = *hi;
  if (first < P) *hi = multisect_point(l, h, P, first);
  if (first > 0) *lo = multisect_point(l, h, P, first - 1);
}

// pivot sequences of the twisted factorisation of T - theta:  forward  s_0 = d_0 - theta, s_{i+1} = d_{i+1} - theta - e_i^2 / s_i
// (dir = +1, written to piv[0..n)), backward p_{n-1} = d_{n-1} - theta, p_i = d_i - theta - e_i^2 / p_{i+1} (dir = -1).  A zero pivot
// is replaced by `tiny` (a perturbation below the rounding error of theta).
TN_HD inline void twisted_pivots(int n, const double* d, const double* e, double theta, int dir, double tiny, double* piv) {
  if (dir > 0) {
    double s = d[0] - theta;
    for (int i = 0; i < n; ++i) {
      if (fabs(s) < tiny) s = s < 0.0 ? -tiny : tiny;
      piv[i] = s;
      if (i + 1 < n) s = (d[i + 1] - theta) - e[i] * (e[i] / s);
    }
  } else {
    double q = d[n - 1] - theta;
    for (int i = n - 1; i >= 0; --i) {
      if (fabs(q) < tiny) q = q < 0.0 ? -tiny : tiny;
      piv[i] = q;
      if (i > 0) q = (d[i - 1] - theta) - e[i - 1] * (e[i - 1] / q);
    }
  }
}

// eigenvector of theta from both pivot sequences: twist at k = argmin |s_k + p_k - (d_k - theta)|, z_k = 1,
// z_i = -(e_i / s_i) z_{i+1} above it and z_{i+1} = -(e_i / p_{i+1}) z_i below; returns the twist index.  z is not normalised.
TN_HD inline int twisted_vector(int n, const double* d, const double* e, double theta, const double* s, const double* pb, double* z) {
  int k = 0;
  double best = fabs(s[0] + pb[0] - (d[0] - theta));
  for (int i = 1; i < n; ++i) {
    const double g = fabs(s[i] + pb[i] - (d[i] - theta));
    if (g < best) { best = g; k = i; }
  }
  z[k] = 1.0;
  for (int i = k - 1; i >= 0; --i) z[i] = -(e[i] / s[i]) * z[i + 1];
  for (int i = k; i + 1 < n; ++i) z[i + 1] = -(e[i] / pb[i + 1]) * z[i];
  return k;
}

}  // namespace tn
