"""Import shim: the product package lives in ``tfswa-unet_b200/`` (the name the build
contract fixes; a hyphen is not importable), this module makes it importable as
``tfswa_unet_b200`` by pointing ``__path__`` at that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tfswa-unet_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
