"""Drop-in for the reference's `models/team29_FreqFusion/__init__.py` (`from .io import main`): test.py does
`from models.team29_FreqFusion import main as model_func` (reference test.py:19-26)."""
import os as _os
import sys as _sys

_ROOT = _os.path.abspath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", ".."))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from isr2_b200.io import main, MODEL_CONFIG  # noqa: E402,F401
