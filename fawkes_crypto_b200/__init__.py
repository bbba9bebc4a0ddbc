"""Importable alias of the `fawkes-crypto_b200/` package directory (a hyphen is not a
valid Python identifier).  All code lives in ../fawkes-crypto_b200/."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "fawkes-crypto_b200"))
from ._pkg import *  # noqa: F401,F403,E402
