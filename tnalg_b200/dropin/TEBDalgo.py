"""Flat drop-in module: `import TEBDalgo` resolves to the CUDA-backed implementation (see tnalg_b200/dropin/MPSClass.py)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from tnalg_b200.TEBDalgo import *  # noqa: E402,F401,F403
from tnalg_b200.TEBDalgo import tebd_standard  # noqa: E402,F401
