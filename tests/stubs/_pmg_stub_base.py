"""TEST-ONLY stand-ins for GUI / plotting / explainability packages the reference's CLI modules import at module level
(test_PMG.py:8-31 → utils/tsne_utils.py, utils/camera_utils.py, utils/benchmark_utils.py, Functions/PMG_evaluation.py) and
that are not installed in this image: matplotlib, tkinter, pytorch_grad_cam, screeninfo, pykalman.  They exist so that the
UNMODIFIED reference scripts can be executed through ``python -m pmg_b200.compat`` up to and through ``model(x)``;
nothing here computes anything.  Never on the product path (tests/stubs is only put on PYTHONPATH by tests)."""


class Anything:
    """Absorbs any attribute access, call, item access, iteration (empty), context-manager use."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return Anything()

    def __call__(self, *a, **k):
        return Anything()

    def __getitem__(self, k):
        return Anything()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __bool__(self):
        return False


def module_getattr(modname):
    def __getattr__(name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(f"module {modname!r} has no attribute {name!r}")
        return Anything()
    return __getattr__
