"""Import helper: the package directory name carries hyphens, so it cannot be named in an
``import`` statement.  ``from b200q_pkg import pkg`` gives the package module."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)

PACKAGE_NAME = "fused-4-bit-dequantize-linear-cuda-kernel_b200"
pkg = importlib.import_module(PACKAGE_NAME)
