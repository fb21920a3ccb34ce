"""Drop-in import name: `from kindel import kindel` resolves to the B200 engine (kindel_b200)."""
import sys as _sys

from kindel_b200 import __version__  # noqa: F401
from kindel_b200 import cli, kindel  # noqa: F401

_sys.modules[__name__ + ".kindel"] = kindel
_sys.modules[__name__ + ".cli"] = cli
