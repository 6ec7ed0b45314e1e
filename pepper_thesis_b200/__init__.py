"""Import alias: the package lives in ``pepper-thesis_b200/`` (not a valid identifier); ``import pepper_thesis_b200``
resolves to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pepper-thesis_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
