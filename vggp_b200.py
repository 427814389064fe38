"""Importable alias of the package directory (its name contains hyphens): `import vggp_b200 as vg`."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
PACKAGE_NAME = "variational-gridded-gaussian-processes_b200"
_pkg = importlib.import_module(PACKAGE_NAME)
sys.modules[__name__] = _pkg
