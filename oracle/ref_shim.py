"""Run the UNMODIFIED reference (/root/reference, root-level generation) in-process.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) validate oracle/dmrg_oracle.py against the real
reference and (b) generate the golden vectors committed under tests/golden/.  It only works where
/root/reference is mounted (the build container); nothing on the GPU box may import it.

The shim is the one documented in SURVEY.md section 8c / Appendix A:
  1. stub modules for imports the container lacks (termcolor, ipdb, matplotlib, mpl_toolkits);
  2. numpy-2 aliases np.complex / np.mat used by HamiltonianModule.py:13,28 and Parameters.py:175...;
  3. scipy.sparse.linalg.eigsh wrapper that flattens the (n,1) v0 the reference passes (MPSClass.py:804),
     installed BEFORE `import MPSClass` because of `from ... import eigsh as eigs` (MPSClass.py:4).
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get('TNALG_REFERENCE_ROOT', '/root/reference')

stats = {'calls': 0, 'matvec': 0}
_installed = False


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'MPSClass.py'))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Make `import MPSClass`, `import DMRG_anyH`, `import Parameters` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    import numpy as np
    sys.dont_write_bytecode = True
    warnings.filterwarnings('ignore', category=SyntaxWarning)
    if 'termcolor' not in sys.modules:
        try:
            import termcolor  # noqa: F401
        except ImportError:
            _stub('termcolor', cprint=lambda *a, **k: None, colored=lambda s, *a, **k: s)
    try:
        import ipdb  # noqa: F401
    except ImportError:
        _stub('ipdb', set_trace=lambda *a, **k: None)
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        mpl = _stub('matplotlib')
        mpl.pyplot = _stub('matplotlib.pyplot')
        mpl.cm = _stub('matplotlib.cm')
        tk = _stub('mpl_toolkits')
        tk.mplot3d = _stub('mpl_toolkits.mplot3d', Axes3D=object)
    if not hasattr(np, 'complex'):
        np.complex = complex
    if not hasattr(np, 'mat'):
        np.mat = np.asmatrix
    import scipy.sparse.linalg as sla
    _eigsh = sla.eigsh

    def eigsh(A, k=6, **kw):
        if kw.get('v0') is not None:
            kw['v0'] = np.asarray(kw['v0']).reshape(-1)
        stats['calls'] += 1
        if isinstance(A, sla.LinearOperator):
            mv = A.matvec

            def counted(x):
                stats['matvec'] += 1
                return mv(x)
            A = sla.LinearOperator(A.shape, matvec=counted, dtype=float)
        return _eigsh(A, k=k, **kw)
    sla.eigsh = eigsh
    sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def modules():
    """Return the reference's (Parameters, DMRG_anyH, MPSClass, TensorBasicModule, HamiltonianModule)."""
    install()
    import importlib
    for name in ('Parameters', 'DMRG_anyH', 'MPSClass', 'TensorBasicModule', 'HamiltonianModule'):
        m = sys.modules.get(name)
        if m is not None and not getattr(m, '__file__', '').startswith(REFERENCE_ROOT):
            raise RuntimeError('module %s already imported from %s (not the reference)' % (name, m.__file__))
    return tuple(importlib.import_module(n) for n in
                 ('Parameters', 'DMRG_anyH', 'MPSClass', 'TensorBasicModule', 'HamiltonianModule'))


def run_finite_dmrg(para, seed=0, quiet=True):
    """np.random.seed(seed); dmrg_finite_size(para) -> (ob, A, info, para)  (DMRG_anyH.py:18-101)."""
    import contextlib
    import io
    import numpy as np
    _, dm, _, _, _ = modules()
    np.random.seed(seed)
    if quiet:
        with contextlib.redirect_stdout(io.StringIO()):
            return dm.dmrg_finite_size(para)
    return dm.dmrg_finite_size(para)
