"""B200-native PMG (PatchGAN-MultiTasks-Gram) eval-forward hot path.

Importable as ``pmg_b200`` (alias package at the repo root) because this directory's name, fixed by
the project layout, contains a hyphen.

    from pmg_b200 import MultiTaskPatchGANGramModelNonOverlapV2     # reference-compatible nn.Module
    from pmg_b200 import install_as_reference_models                # bind the reference's own utils to it

The CUDA library (``libpmg_b200.so``, built in-tree by ``pmg_b200.build``) is mandatory: importing this
package without it raises ImportError.  There is no CPU or eager-PyTorch fallback.
"""
from . import _lib                      # noqa: F401  (fails loudly if the .so is missing)
from . import models_PMG                # noqa: F401
from . import models_PM                 # noqa: F401
from .models_PMG import MultiTaskPatchGANGramModelNonOverlapV2  # noqa: F401
from .models_PM import MultiTaskPatchGAN  # noqa: F401
from .compat import install_as_reference_models  # noqa: F401
from . import gram as gram              # noqa: F401
from . import runner as runner          # noqa: F401
from . import prepost as prepost        # noqa: F401

__all__ = ["MultiTaskPatchGANGramModelNonOverlapV2", "MultiTaskPatchGAN", "install_as_reference_models", "gram", "runner", "prepost"]
