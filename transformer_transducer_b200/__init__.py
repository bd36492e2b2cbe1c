"""Import shim: the package sources live in ``transformer-transducer_b200/`` (the name the build
contract asks for, which is not a Python identifier); this module makes them importable as
``transformer_transducer_b200`` without symlinks."""
import os as _os

_src = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "transformer-transducer_b200")
if not _os.path.isdir(_src):
    raise ImportError("transformer-transducer_b200/ not found next to %s" % __file__)
__path__.append(_src)

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
