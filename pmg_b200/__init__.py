"""Importable alias of the package directory
``heuristic_style_transfer_for_real-time_efficient_weather_attribute_detection_b200`` (whose mandated
name contains a hyphen and therefore cannot appear in an ``import`` statement)."""
import importlib
import os
import sys

_REAL_NAME = "heuristic_style_transfer_for_real-time_efficient_weather_attribute_detection_b200"
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_real = importlib.import_module(_REAL_NAME)
for _k, _m in list(sys.modules.items()):
    if _k == _REAL_NAME or _k.startswith(_REAL_NAME + "."):
        sys.modules[__name__ + _k[len(_REAL_NAME):]] = _m
