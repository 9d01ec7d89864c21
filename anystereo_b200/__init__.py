"""Importable alias: the product package lives in the directory ``any-stereo_b200/`` (the name the task
fixes), which is not a valid Python identifier.  ``import anystereo_b200`` resolves to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "any-stereo_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
