"""Import shim: the package directory is `ws-mgmap_b200/` (not a valid Python identifier).
`import wsmgmap_b200` loads that directory as the package *named* wsmgmap_b200, so every
submodule exists exactly once (wsmgmap_b200.ops, wsmgmap_b200._lib, ...)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ws-mgmap_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _pkg
_spec.loader.exec_module(_pkg)
