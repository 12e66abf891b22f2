"""TEST INFRASTRUCTURE ONLY: the plain-C restatement (oracle/liboracle.so).

See oracle/kaori_port.c.  Never imported by screencounter_b200/.
"""
import os
import sys

from ._binding import OracleBinding, KaoriError, DUP_FIRST, DUP_LAST, DUP_NONE, DUP_ERROR  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_B = OracleBinding(os.path.join(_HERE, "liboracle.so"), "kport_", "run `make -C oracle liboracle.so`")

available = _B.available
for _name in dir(_B):
    if not _name.startswith("_") and _name not in ("available", "path", "prefix", "hint"):
        setattr(sys.modules[__name__], _name, getattr(_B, _name))
