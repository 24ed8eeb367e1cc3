"""``python -m pmg_b200.build [--force] [-v]`` — runnable shim of the real package's in-tree nvcc build."""
import os as _os
import sys as _sys
import importlib.util as _u

# loaded by file path: building must work BEFORE the library exists (importing the package dlopens it)
_p = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                   "heuristic_style_transfer_for_real-time_efficient_weather_attribute_detection_b200", "build.py")
_spec = _u.spec_from_file_location("_pmg_b200_build", _p)
_real = _u.module_from_spec(_spec)
_spec.loader.exec_module(_real)
build = _real.build

if __name__ == "__main__":
    print(build(force="--force" in _sys.argv, verbose="-v" in _sys.argv))
