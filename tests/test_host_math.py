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
