"""Import shim: `import sscvae` loads the package directory `style-seqcvae_b200/` (whose name is not
a valid Python identifier) under the module name `style_seqcvae_b200` and re-exports its API."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "style-seqcvae_b200")
_NAME = "style_seqcvae_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]

from style_seqcvae_b200 import *          # noqa: E402,F401,F403
from style_seqcvae_b200 import _lib       # noqa: E402,F401
__all__ = _pkg.__all__
