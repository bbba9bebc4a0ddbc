"""Package body of fawkes_crypto_b200 (see ../fawkes_crypto_b200/__init__.py)."""
from . import native  # noqa: F401
from .groth16 import (Parameters, Proof, VK, G1Point, G2Point, Circuit,  # noqa: F401
                      setup, prove, prove_with_rs, prove_batch, ProveStream, verify, Context)
