"""Importable alias of the product package, whose directory name (`image-super-resolution-2_b200/`)
is not a valid Python identifier.  `import isr2_b200` executes that directory's __init__.py and
resolves sub-modules (`isr2_b200.io`, `isr2_b200.hat`, ...) from it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "image-super-resolution-2_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
