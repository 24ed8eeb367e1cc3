"""TEST-ONLY stub of `pytorch_grad_cam.utils` (not installed in this image) — see tests/stubs/_pmg_stub_base.py."""
from _pmg_stub_base import Anything, module_getattr

__getattr__ = module_getattr("pytorch_grad_cam.utils")
