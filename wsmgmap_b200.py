"""Import alias: the package directory is `ws-mgmap_b200/` (not a valid Python
identifier), so `import wsmgmap_b200` resolves to it."""
import importlib
import sys

_pkg = importlib.import_module("ws-mgmap_b200")
sys.modules[__name__] = _pkg
