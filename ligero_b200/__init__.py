"""ligero_b200 -- B200-native back end for the Ligero commit-and-test hot path (NP-Eng/ligero).

The numeric work is hand-written sm_100a CUDA behind a C ABI (include/ligero_b200.h); this package is
the host-side mirror used by tests and benchmarks.  There is no CPU fallback: creating a Context
without the built library or without a GPU raises LigeroB200Error.
"""
from ._lib import LIB_PATH, LigeroB200Error  # noqa: F401
from .backend import BN254_R, CommittedMatrix, Constraints, Context, fr_to_limbs, limbs_to_fr  # noqa: F401
from .api import (ArithmeticCircuit, LigeroCircuit, LigeroProof, PoseidonSponge, DEFAULT_SECURITY_LEVEL,  # noqa: F401,E402
                  CHACHA_SEED_BYTES)
