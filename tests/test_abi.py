"""The C-ABI library loads without a GPU and exports every symbol include/tnalg_b200.h declares (no compute call)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'tnalg_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tn_[a-z0-9_]+)\s*\(', text)) - {'tn_allreduce_fn', 'tn_matvec_fn'})


def test_library_exports_every_declared_symbol():
    from tnalg_b200 import _lib, build
    build.build(verbose=False)
    lib = _lib.load_library()
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.SIGNATURES.keys())
    assert lib.tn_version() >= 100
    assert lib.tn_chain_gemm_workspace_bytes(1, 1) > 0 and lib.tn_svd_workspace_bytes(8, 4) > 0


def test_product_fails_loudly_without_cuda():
    import torch
    from tnalg_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.TnError):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'tnalg_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(base, f)).read()
                assert 'oracle' not in src.replace('oracle-backed', ''), os.path.join(base, f)


def test_bad_arguments_are_errors_not_ub():
    """shape / null-pointer mistakes return a status and a message (checked before any device work)"""
    import ctypes as C
    from tnalg_b200 import _lib
    lib = _lib.load_library()
    st = lib.tn_env_update(5, None, 0, 2, 0, 0, None, None, None, None, None, None, 0, None)
    assert st == -1 and b'tn_env_update' in lib.tn_last_error()
    st = lib.tn_svd_jacobi(None, 0, 0, 0, None, None, None, None, None, 0, None)
    assert st == -1
    handle = C.c_void_p()
    st = lib.tn_effh_plan_create(C.byref(handle), 4, 10, 4, None, None, None, 0, None, None, 0, None, None, 0, None, None,
                                 None, 0, 1, None, 0, None)
    assert st == -1 and b'bad shape' in lib.tn_last_error()
    st = lib.tn_qr_householder(None, 0, 0, 0, None, 0, None, None, 0, None)
    assert st == -1 and b'tn_qr_householder' in lib.tn_last_error()
    st = lib.tn_comm_init(None, None, 0, 1)
    assert st == -1
    st = lib.tn_ed_apply(None, None, 4, 2, 0, None, None, None, None, 0, 0.0, 1.0, None, 0, None)
    assert st == -1
    assert lib.tn_qr_workspace_bytes(64, 32) > 0 and lib.tn_eigh_workspace_bytes(16) > 0 and lib.tn_ed_workspace_bytes(4, 2, 3, 1, 20) > 0
    assert lib.tn_set_deterministic(1) == 0 and lib.tn_set_deterministic(0) == 1
