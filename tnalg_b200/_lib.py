"""ctypes binding of libtnalg_b200.so (the C ABI declared in include/tnalg_b200.h).

There is no CPU fallback: `load()` raises when the shared library is missing or when CUDA is unavailable.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TNALG_B200_LIB', os.path.join(HERE, 'libtnalg_b200.so'))  # override: kernel-variant experiments

MAX_D = 4          # TN_MAX_LOADPATH_DIM: site operators applied inside the GEMM operand path (tn_link.op)
MAX_PHYS_D = 9     # TN_MAX_PHYS_DIM: effective-Hamiltonian plans and the element-wise site-operator kernel

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_void_pp = C.POINTER(C.c_void_p)


class TnLink(C.Structure):
    _fields_ = [('A', C.c_void_p), ('B', C.c_void_p), ('op', C.c_double * (MAX_D * MAX_D)), ('has_op', C.c_int),
                ('reserved', C.c_int)]


class TnProblem(C.Structure):
    _fields_ = [('C', C.c_void_p), ('alpha', C.c_double), ('link_begin', C.c_int), ('link_count', C.c_int),
                ('accumulate', C.c_int), ('reserved', C.c_int)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p)
MATVEC_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
c_longlong_p = C.POINTER(C.c_longlong)

# name -> (restype, argtypes); must list every symbol of include/tnalg_b200.h (tests/test_abi.py checks that)
SIGNATURES = {
    'tn_last_error': (C.c_char_p, []),
    'tn_version': (C.c_int, []),
    'tn_device_info': (C.c_int, [c_int_p, c_int_p, c_int_p]),
    'tn_launch_count': (C.c_longlong, []),
    'tn_launch_count_reset': (None, []),
    'tn_set_deterministic': (C.c_int, [C.c_int]),
    'tn_measure_dmma_peak': (C.c_int, [c_double_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_chain_gemm_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'tn_chain_gemm': (C.c_int, [C.c_int] * 8 + [C.POINTER(TnProblem), C.c_int, C.POINTER(TnLink), C.c_int, C.c_int,
                                              C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_env_update_workspace_bytes': (C.c_size_t, [C.c_int] * 5),
    'tn_env_update': (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_void_pp, c_int_p, c_void_pp,
                                c_double_p, c_int_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_lincomb': (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, c_void_pp, c_double_p, C.c_void_p]),
    'tn_apply_site_op': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, C.c_void_p]),
    'tn_dot': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_dot_workspace_bytes': (C.c_size_t, [C.c_longlong]),
    'tn_trace': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'tn_effh_plan_workspace_bytes': (C.c_size_t, [C.c_int] * 6),
    'tn_effh_plan_create': (C.c_int, [c_void_pp, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, c_double_p,
                                      C.c_int, c_void_pp, c_double_p, C.c_int, c_void_pp, c_double_p,
                                      C.c_int, c_void_pp, c_void_pp, c_double_p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_effh_plan_create_rows': (C.c_int, [c_void_pp, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, c_double_p,
                                           C.c_int, c_void_pp, c_double_p, C.c_int, c_void_pp, c_double_p,
                                           C.c_int, c_void_pp, c_void_pp, c_double_p, C.c_int, C.c_int,
                                           C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_effh_plan_rows': (C.c_int, [C.c_void_p, c_int_p, c_int_p]),
    'tn_effh_matvec': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]),
    'tn_effh_plan_flops': (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    'tn_effh_plan_uses_tma': (C.c_int, [C.c_void_p]),
    'tn_effh_plan_destroy': (C.c_int, [C.c_void_p]),
    'tn_lanczos_workspace_bytes': (C.c_size_t, [C.c_longlong, C.c_int]),
    'tn_lanczos_workspace_bytes_sharded': (C.c_size_t, [C.c_void_p, C.c_void_p, C.c_int]),
    'tn_lanczos_lm1': (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_int, C.c_int, c_double_p,
                                 C.c_void_p, c_int_p, c_double_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                 C.c_void_p]),
    'tn_lanczos_generic': (C.c_int, [MATVEC_FN, C.c_void_p, C.c_longlong, C.c_double, C.c_void_p, C.c_double, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int, C.c_longlong, c_double_p, C.c_void_p, c_int_p, c_double_p, C.c_void_p,
                                     C.c_size_t, C.c_void_p]),
    'tn_comm_unique_id': (C.c_int, [C.c_char_p]),
    'tn_comm_init_rank': (C.c_int, [c_void_pp, C.c_char_p, C.c_int, C.c_int]),
    'tn_comm_init': (C.c_int, [c_void_pp, C.c_void_p, C.c_int, C.c_int]),
    'tn_comm_rank': (C.c_int, [C.c_void_p]),
    'tn_comm_world': (C.c_int, [C.c_void_p]),
    'tn_comm_collectives': (C.c_longlong, [C.c_void_p]),
    'tn_comm_peer_collectives': (C.c_longlong, [C.c_void_p]),
    'tn_comm_destroy': (C.c_int, [C.c_void_p]),
    'tn_comm_allreduce_sum': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    'tn_comm_allgather': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    'tn_comm_broadcast_many': (C.c_int, [C.c_void_p, c_void_pp, c_longlong_p, c_int_p, C.c_int, C.c_void_p]),
    'tn_expect_workspace_bytes': (C.c_size_t, [c_int_p, C.c_int, C.c_int, C.c_int]),
    'tn_expect_1body': (C.c_int, [c_void_pp, c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, c_double_p, c_double_p, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    'tn_expect_2body': (C.c_int, [c_void_pp, c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_double_p, c_double_p,
                                  c_double_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_eigh_workspace_bytes': (C.c_size_t, [C.c_int]),
    'tn_eigh_jacobi': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, c_int_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_qr_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'tn_qr_householder': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.c_void_p]),
    'tn_qr_l2r': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_qr_r2l': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_ed_workspace_bytes': (C.c_size_t, [C.c_int] * 5),
    'tn_ed_apply': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_double_p, C.c_int,
                              C.c_double, C.c_double, C.c_void_p, C.c_size_t, C.c_void_p]),
    'tn_ed_ground_state': (C.c_int, [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_double_p, C.c_int, C.c_double, C.c_void_p,
                                     C.c_double, C.c_int, C.c_int, c_double_p, C.c_void_p, c_int_p, c_double_p, C.c_void_p, C.c_size_t,
                                     C.c_void_p]),
    'tn_svd_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'tn_svd_jacobi': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, c_int_p,
                                C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


class TnError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    """dlopen the shared library and attach the signatures (works without a GPU; no compute call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(path):
        raise TnError('libtnalg_b200.so not found at %s -- build it with `python -m tnalg_b200.build` '
                      '(there is no CPU fallback)' % path)
    import torch  # noqa: F401  (loads libcudart.so.12 so that the library resolves the same runtime)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def load():
    """library + device check; raises when no sm_100 GPU is visible."""
    import torch
    lib = load_library()
    if not torch.cuda.is_available():
        raise TnError('tnalg_b200 needs a CUDA device (sm_100a); torch.cuda.is_available() is False and there is '
                      'no CPU fallback')
    return lib


def check(status, lib=None):
    if status != 0:
        lib = lib or _lib
        msg = lib.tn_last_error().decode() if lib is not None else ''
        raise TnError('tnalg_b200 error %d: %s' % (status, msg))
