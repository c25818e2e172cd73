"""tnalg_b200/csrc/host_math.h (tridiagonal QL used by the Lanczos Ritz kernel, the Jacobi rotation of the SVD kernels)
compiled with g++ and checked against numpy on the CPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "host_math.h"
extern "C" int t_ql(int n, double* d, double* e, double* z) { return tn::tridiag_ql(n, d, e, z, n); }
extern "C" int t_ql_rows(int n, double* d, double* e, double* z, int k0, int ks) { return tn::tridiag_ql_rows(n, d, e, z, n, k0, ks); }
extern "C" void t_rot(double a, double b, double g, double* c, double* s) { tn::jacobi_rotation(a, b, g, c, s); }
extern "C" int t_count(int n, const double* d, const double* e2, double x) { return tn::tridiag_count_below(n, d, e2, x); }
// serial emulation of the fast path of lanczos_ritz_kernel (the lanes of the warp become a loop): which = 0 smallest / 1 largest
// eigenvalue; returns the residual max|T u - theta u| of the normalised matrix
extern "C" double t_extreme(int n, const double* d, const double* e, int which, double* theta, double* u) {
  using namespace tn;
  double c, r;
  tridiag_gershgorin(n, d, e, &c, &r);
  double sd[kMaxNcv], se[kMaxNcv], se2[kMaxNcv], pf[kMaxNcv], pb[kMaxNcv];
  for (int i = 0; i < n; ++i) { sd[i] = (d[i] - c) / r; se[i] = i + 1 < n ? e[i] / r : 0.0; se2[i] = se[i] * se[i]; }
  const int k = which ? n - 1 : 0;
  double lo = 1e300, hi = -1e300;
  for (int i = 0; i < n; ++i) {
    const double rr = (i > 0 ? fabs(se[i - 1]) : 0.0) + fabs(se[i]);
    lo = fmin(lo, sd[i] - rr);
    hi = fmax(hi, sd[i] + rr);
  }
  lo -= 0x1p-40;
  hi += 0x1p-40;
  for (int round = 0; round < 48; ++round) {
    int first = 16;
    for (int idx = 15; idx >= 0; --idx)
      if (tridiag_count_below(n, sd, se2, multisect_point(lo, hi, 16, idx)) >= k + 1) first = idx;
    multisect_shrink(&lo, &hi, 16, first);
    if (!((hi - lo) > 2.5e-16) || !(multisect_point(lo, hi, 16, 0) > lo)) break;
  }
  const double xs = 0.5 * (lo + hi);
  *theta = c + r * xs;
  twisted_pivots(n, sd, se, xs, +1, 1e-280, pf);
  twisted_pivots(n, sd, se, xs, -1, 1e-280, pb);
  twisted_vector(n, sd, se, xs, pf, pb, u);
  double n2 = 0.0;
  for (int i = 0; i < n; ++i) n2 += u[i] * u[i];
  const double inv = 1.0 / sqrt(n2);
  double rmax = 0.0;
  for (int i = 0; i < n; ++i) {
    const double res = (sd[i] - xs) * u[i] + (i > 0 ? se[i - 1] * u[i - 1] : 0.0) + (i + 1 < n ? se[i] * u[i + 1] : 0.0);
    rmax = fmax(rmax, fabs(res) * inv);
  }
  for (int i = 0; i < n; ++i) u[i] *= inv;
  return rmax;
}
'''


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp('host_math')
    src, so = d / 'hm.cpp', d / 'hm.so'
    src.write_text(SRC)
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-I', os.path.join(ROOT, 'tnalg_b200', 'csrc'), str(src), '-o', str(so)])
    L = ctypes.CDLL(str(so))
    dp = ctypes.POINTER(ctypes.c_double)
    L.t_ql.argtypes = [ctypes.c_int, dp, dp, dp]
    L.t_ql_rows.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int, ctypes.c_int]
    L.t_rot.argtypes = [ctypes.c_double] * 3 + [dp, dp]
    L.t_rot.restype = None
    L.t_count.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double]
    L.t_extreme.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, dp, dp]
    L.t_extreme.restype = ctypes.c_double
    return L


def _p(x):
    return x.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@pytest.mark.parametrize('n', [1, 2, 3, 8, 20, 64])
def test_tridiag_ql_vs_numpy(lib, n):
    rng = np.random.RandomState(n)
    d, e = rng.randn(n), np.append(rng.randn(n - 1), 0.0)
    T = np.diag(d) + np.diag(e[:-1], 1) + np.diag(e[:-1], -1)
    dd, ee, z = d.copy(), e.copy(), np.eye(n)
    assert lib.t_ql(n, _p(dd), _p(ee), _p(z)) == 0
    w = np.linalg.eigvalsh(T)
    assert np.abs(np.sort(dd) - w).max() < 1e-13 * max(1.0, np.abs(w).max())
    assert np.abs(T @ z - z * dd[None, :]).max() < 1e-13 * max(1.0, np.abs(w).max())
    assert np.abs(z.T @ z - np.eye(n)).max() < 1e-13
    # row-sliced variant: every "lane" updates its own rows of z and all get the same eigenvalues
    z2 = np.eye(n)
    for k0 in range(4):
        d2, e2 = d.copy(), e.copy()
        assert lib.t_ql_rows(n, _p(d2), _p(e2), _p(z2), k0, 4) == 0
        assert np.array_equal(d2, dd)
    assert np.array_equal(z2, z)


def test_jacobi_rotation_orthogonalises_pairs(lib):
    rng = np.random.RandomState(0)
    c, s = ctypes.c_double(), ctypes.c_double()
    for scale_p, scale_q in [(1, 1), (1, 1e-9), (1e-12, 1), (1e-150, 1e-150), (1e140, 1e130), (1, 1e-17)]:
        for _ in range(20):
            p, q = scale_p * rng.randn(40), scale_q * rng.randn(40)
            a, b, g = p @ p, q @ q, p @ q
            if g == 0.0:
                continue
            lib.t_rot(a, b, g, ctypes.byref(c), ctypes.byref(s))
            assert abs(c.value ** 2 + s.value ** 2 - 1) < 1e-15
            p2, q2 = c.value * p - s.value * q, s.value * p + c.value * q
            assert abs(p2 @ q2) <= 1e-15 * np.linalg.norm(p2) * np.linalg.norm(q2) + 1e-300, (scale_p, scale_q)
            # the small rotation is chosen (|t| <= 1): the larger row stays the larger one
            assert abs(s.value) <= abs(c.value) + 1e-16


def _lanczos_like(n, rng, decay):
    """tridiagonal of a converging Lanczos run: O(1) entries first, couplings decaying to `decay` at the end"""
    d = rng.randn(n)
    e = np.abs(rng.randn(n)) * np.logspace(0, np.log10(decay), n)
    return d, e


@pytest.mark.parametrize('n', [2, 3, 5, 20, 21, 40, 64])
def test_sturm_counts_and_extreme_pair_vs_numpy(lib, n):
    rng = np.random.RandomState(100 + n)
    cases = [(rng.randn(n), rng.randn(n)), _lanczos_like(n, rng, 1e-9), _lanczos_like(n, rng, 1e-200),
             (np.full(n, 0.3), np.full(n, 1e-12)),                       # nearly a multiple of the identity
             (1e8 * rng.randn(n), 1e8 * rng.randn(n)), (1e-8 * rng.randn(n), 1e-8 * rng.randn(n)),
             (np.arange(n, dtype=float), np.zeros(n)),                   # diagonal (fully reducible)
             (np.r_[rng.randn(n // 2), 5 + rng.randn(n - n // 2)], np.r_[rng.randn(n // 2 - 1), 0.0, rng.randn(n - n // 2)])]  # two blocks
    for d, e in cases:
        d, e = np.ascontiguousarray(d, dtype=float), np.ascontiguousarray(e, dtype=float)
        T = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
        w, v = np.linalg.eigh(T)
        scale = max(np.abs(w).max(), 1e-300)
        e2 = e * e
        for x in np.r_[w[0] - 1.0 * scale, 0.5 * (w[:-1] + w[1:]), w[-1] + 1.0 * scale]:
            gaps = np.abs(w - x).min()
            if gaps > 1e-12 * scale:
                assert lib.t_count(n, _p(d), _p(e2), float(x)) == int((w < x).sum())
        for which in (0, 1):
            theta, u = ctypes.c_double(), np.zeros(n)
            resid = lib.t_extreme(n, _p(d), _p(e), which, ctypes.byref(theta), _p(u))
            ref = w[-1] if which else w[0]
            assert abs(theta.value - ref) <= 2e-15 * scale + 1e-300, (n, which, theta.value, ref)
            assert resid <= 2e-14, (n, which, resid)
            assert abs(np.linalg.norm(u) - 1) < 1e-14
            assert np.abs(T @ u - theta.value * u).max() <= 1e-13 * scale
            sep = (w[-1] - w[-2]) if which else (w[1] - w[0])
            if sep > 1e-6 * scale:      # well separated: the vector itself is determined
                vref = v[:, -1] if which else v[:, 0]
                assert abs(abs(u @ vref) - 1) < 1e-9
