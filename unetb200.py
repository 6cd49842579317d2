"""Import shim: the product package lives in `semantic-segmentation-unet_b200/` (not a valid Python identifier);
`import unetb200` loads it under this name so `unetb200.model`, `unetb200._C`, ... resolve normally."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "semantic-segmentation-unet_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
