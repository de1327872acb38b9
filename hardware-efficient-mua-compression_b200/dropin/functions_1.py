"""`from functions_1 import *` shim: put this directory on sys.path in place of the reference's
`Compressing data/` and the three driver scripts pick up the B200 implementations."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
import mua_b200  # noqa: E402,F401
from mua_b200.functions_1 import *  # noqa: E402,F401,F403
from mua_b200.functions_1 import __all__  # noqa: E402,F401
