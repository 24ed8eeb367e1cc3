"""TEST-ONLY stub of `matplotlib.pyplot` (not installed in this image) — see tests/stubs/_pmg_stub_base.py."""
from _pmg_stub_base import Anything, module_getattr

__getattr__ = module_getattr("matplotlib.pyplot")
