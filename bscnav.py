"""Import alias: `import bscnav` == the package in ./bsc-nav_b200 (whose directory name is not a Python identifier)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("bsc-nav_b200")
