"""Binding the UNMODIFIED reference scripts to the B200 module.

The reference imports its model by module path (``from Models.models_PMG import
MultiTaskPatchGANGramModelNonOverlapV2`` — utils/PMG/PMG_model_utils.py:9-11) and is run as
``python test_PMG.py …`` from its own checkout, so ``sys.path[0]`` always resolves ``Models`` to the
reference's package.  ``install_as_reference_models()`` pre-seeds ``sys.modules['Models.models_PMG']``
with this package's module, after which ``build_model_from_hparams`` / ``load_best_model`` /
``test_PMG.py --mode benchmark`` / ``Functions/PMG_evaluation.py`` run unchanged on the CUDA kernels.

    python -m pmg_b200.compat /path/to/reference/test_PMG.py --mode benchmark …
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types


def install_as_reference_models() -> None:
    from . import models_PM, models_PMG
    pkg = sys.modules.get("Models")
    if pkg is None:
        try:
            pkg = importlib.import_module("Models")       # the reference's namespace/package if on sys.path
        except ImportError:
            pkg = types.ModuleType("Models")
            pkg.__path__ = []                              # mark as package
            sys.modules["Models"] = pkg
    sys.modules["Models.models_PMG"] = models_PMG
    setattr(pkg, "models_PMG", models_PMG)
    sys.modules["Models.models_PM"] = models_PM            # test_PM.py:16-21 imports MultiTaskPatchGAN & helpers
    setattr(pkg, "models_PM", models_PM)
    bind_reference_pm_helpers()


_REF_PM_NAME = "_pmg_b200_reference_models_PM"


def bind_reference_pm_helpers() -> bool:
    """test_PM.py:16-21 imports ``TaskSpecificModel, load_model_weights, checkpoint_has_se, print_model_parameters`` from
    ``Models.models_PM`` next to the model class.  They are host-side helpers (checkpoint loading, parameter printing,
    a one-task wrapper), not part of the forward path, so the reference's own definitions are used: its
    ``Models/models_PM.py`` is executed under a private module name and the four names are bound onto this package's
    ``models_PM``.  Returns False when no reference checkout is on ``sys.path``."""
    from . import models_PM
    if all(h in models_PM.__dict__ for h in models_PM._REFERENCE_HELPERS):
        return True
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    roots = list(getattr(sys.modules.get("Models"), "__path__", [])) + [os.path.join(p or ".", "Models") for p in sys.path]
    for d in roots:
        path = os.path.join(d, "models_PM.py")
        if not os.path.isfile(path) or os.path.dirname(os.path.abspath(path)) == here:
            continue
        spec = importlib.util.spec_from_file_location(_REF_PM_NAME, path)
        ref = importlib.util.module_from_spec(spec)
        try:
            spec.loader.exec_module(ref)
        except Exception:
            continue
        if not all(hasattr(ref, h) for h in models_PM._REFERENCE_HELPERS):
            continue
        sys.modules[_REF_PM_NAME] = ref
        for h in models_PM._REFERENCE_HELPERS:
            setattr(models_PM, h, getattr(ref, h))
        return True
    return False


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    sys.path.insert(0, os.path.dirname(script))            # what `python script.py` would have done
    install_as_reference_models()
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
