"""Importable alias for the product package.

The product lives in ``swin-transformer-object-detection_b200/`` (the repository's required
package directory, which is not a legal Python identifier); this stub makes it importable as
``swin_b200`` by pointing the package search path at that directory and running its
``__init__``."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "swin-transformer-object-detection_b200")
__path__ = [_impl]
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
del _os, _f
