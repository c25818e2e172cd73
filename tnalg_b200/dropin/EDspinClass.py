"""Flat drop-in module: `import EDspinClass` resolves to the CUDA-backed implementation (see tnalg_b200/dropin/MPSClass.py)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from tnalg_b200.EDspinClass import *  # noqa: E402,F401,F403
from tnalg_b200.EDspinClass import EDbasic, exact_ground_state  # noqa: E402,F401
