"""``python -m pmg_b200.compat /path/to/reference/test_PMG.py --mode benchmark …`` — runnable shim of the real
package's ``compat`` module (see pmg_b200/__init__.py); re-exports its public names."""
import importlib as _importlib

_real = _importlib.import_module("heuristic_style_transfer_for_real-time_efficient_weather_attribute_detection_b200.compat")
install_as_reference_models = _real.install_as_reference_models
bind_reference_pm_helpers = _real.bind_reference_pm_helpers
main = _real.main
__doc__ = _real.__doc__

if __name__ == "__main__":
    raise SystemExit(main())
