"""Importable alias of the product package.

The product lives in `unsupervised-part-segmentation_b200/` (the directory name the project
layout prescribes); a hyphen cannot appear in a Python module name, so `ups_b200` points its
package search path there and executes that package's __init__.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "unsupervised-part-segmentation_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
