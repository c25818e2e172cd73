"""Flat drop-in module: `import Parameters` resolves here when tnalg_b200/dropin is on sys.path in place of the reference
tree.  Re-exports tnalg_b200.Parameters; classes are re-homed so that pickles record 'Parameters.<Class>' like the reference's."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from tnalg_b200.Parameters import *  # noqa: F401,F403,E402
from tnalg_b200 import Parameters as _impl  # noqa: E402

for _name, _obj in list(vars(_impl).items()):
    if isinstance(_obj, type) and _obj.__module__ == _impl.__name__:
        _obj.__module__ = __name__
        globals()[_name] = _obj
    elif not _name.startswith('__'):
        globals().setdefault(_name, _obj)
