"""Drop-in module: with `convdr_b200/shim` on PYTHONPATH, `import faiss` in ConvDR's
drivers/run_convdr_inference.py resolves to the B200 engine's facade."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from convdr_b200.faiss_compat import *  # noqa: F401,F403,E402
from convdr_b200.faiss_compat import get_num_gpus  # noqa: F401,E402
