"""Importable alias of the product package.

The product lives in ``multimodal-object-detection-via-probabilistic-ensembling_b200/`` (the directory
name the build contract asks for, which is not a valid Python identifier); this shim puts that directory
on the package ``__path__`` so ``import probenb200.fusion`` etc. resolve there.
"""
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
PACKAGE_DIR = _os.path.join(_ROOT, "multimodal-object-detection-via-probabilistic-ensembling_b200")
__path__.append(PACKAGE_DIR)

__version__ = "0.1.0"
