"""Import shim: the product package lives in ``caro-ai_b200/`` (the name the build contract asks
for, which is not a valid Python identifier); this makes it importable as ``caro_ai_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "caro-ai_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
