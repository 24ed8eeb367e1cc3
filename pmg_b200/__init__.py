"""Importable alias of the package directory
``heuristic_style_transfer_for_real-time_efficient_weather_attribute_detection_b200`` (whose mandated
name contains a hyphen and therefore cannot appear in an ``import`` statement).

Every sub-module of the real package is reachable as ``pmg_b200.<name>`` (same module objects).  The two that are
also run as scripts — ``python -m pmg_b200.compat <reference script> …`` and ``python -m pmg_b200.build`` — have
small real files in this directory, because ``runpy`` needs a spec whose name matches the ``-m`` argument."""
import importlib
import os
import sys

_REAL_NAME = "heuristic_style_transfer_for_real-time_efficient_weather_attribute_detection_b200"
_RUNNABLE = ("compat", "build")
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_real = importlib.import_module(_REAL_NAME)
for _k, _m in list(sys.modules.items()):
    if _k == _REAL_NAME:
        continue
    if _k.startswith(_REAL_NAME + ".") and _k[len(_REAL_NAME) + 1:] not in _RUNNABLE:
        sys.modules[__name__ + _k[len(_REAL_NAME):]] = _m
for _n in dir(_real):
    if not _n.startswith("__") and _n not in _RUNNABLE:
        globals()[_n] = getattr(_real, _n)
__all__ = list(getattr(_real, "__all__", []))
__doc__ = (_real.__doc__ or "") + "\n" + __doc__
