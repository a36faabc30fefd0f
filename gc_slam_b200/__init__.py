"""
Import shim: the package sources live in ``gc-slam_b200/`` (the name the build contract fixes),
which is not a valid Python identifier.  This module makes them importable as ``gc_slam_b200``.
"""
import os as _os

_src = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gc-slam_b200")
__path__.insert(0, _src)
with open(_os.path.join(_src, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_src, "__init__.py"), "exec"))
del _os, _f
