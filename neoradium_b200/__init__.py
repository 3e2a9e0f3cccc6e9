"""neoradium_b200 -- B200-native (sm_100a) implementation of NeoRadium's 5G NR LDPC hot path.

Drop-in for ``neoradium.ldpc`` / ``neoradium.chancodebase``:

    from neoradium_b200 import LdpcEncoder, LdpcDecoder, ChanCodeBase

Everything that touches payload bits or LLRs runs in hand-written CUDA kernels behind the C-ABI of include/nrldpc.h
(neoradium_b200/csrc).  Importing this package does not need a GPU; the first compute call does, and fails loudly
without one (no CPU fallback).
"""
from .chancodebase import ChanCodeBase, strToPoly          # noqa: F401
from .ldpc import LdpcBase, LdpcDecoder, LdpcEncoder       # noqa: F401
from . import _native                                       # noqa: F401

__version__ = "0.1.0"


def build(force=False, verbose=False):
    """Compile libnrldpc.so in-tree for sm_100a."""
    return _native.build(force=force, verbose=verbose)
