"""Bare-name shim: put this directory on sys.path and `from trainer import *` resolves to the B200 engine
(the reference imports its modules by bare name, selfplay.py:3-6)."""
import importlib as _il
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_m = _il.import_module("2048nn_b200.trainer")
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
