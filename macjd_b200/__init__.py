"""Importable alias of the product package.

The product lives in ``ma-cjd-cooperative-jamming-decision-making-via-marl_b200/``
(the directory name the build contract prescribes, not a valid Python
identifier); this shim exposes it as ``macjd_b200`` by pointing the package
search path at that directory.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "ma-cjd-cooperative-jamming-decision-making-via-marl_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
