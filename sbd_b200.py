"""Import shim: the package directory is named after the reference repository
(`semi-blind-image-deblurring-problems-with-tv_b200/`), which is not a valid
Python identifier.  `import sbd_b200` loads that directory as the package
`sbd_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "semi-blind-image-deblurring-problems-with-tv_b200")
_spec = importlib.util.spec_from_file_location(
    "sbd_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sbd_b200"] = _mod
_spec.loader.exec_module(_mod)
