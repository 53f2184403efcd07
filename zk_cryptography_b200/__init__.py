"""zk_cryptography_b200 -- host-side mirror of the reference's `polynomial` / `sumcheck` API over the
zksc C ABI (include/zksc.h), whose compute path is hand-written CUDA for sm_100a.

The classes keep the reference's names and argument meaning so that tests read like the reference's:
    Multilinear, ComposedMultilinear                      (polynomial/src/...)
    Sumcheck, ComposedSumcheck,
    MultiComposedSumcheckProver / MultiComposedSumcheckVerifier   (sumcheck/src/...)
    Gate, CircuitLayer, Circuit, GKRProtocol, GKRProof    (circuit/src/..., gkr/src/...: the caller of prove_partial)
    LayeredCircuit                                        (layered circuits of any widths, linear-time GKR prover: BASELINE config 4 as written)
    MultilinearKZG, TrustedSetup, SuccintGKRProtocol      (kzg/src/multilinear_kzg.rs, gkr/src/succint_protocol.rs: commitment / open / prove)
Field elements cross the boundary as ark-ff's in-memory form (4 x u64, little-endian, Montgomery);
on the Python side a table is a numpy array of shape (n, 4), dtype uint64, and scalars are Python
ints (canonical residues).  There is no CPU fallback: without libzksc.so or without a CUDA device the
compute entry points raise.
"""
from ._lib import (ZkscError, Context, Tables, lib, R_MOD, to_mont, from_mont, PROTO_SUMCHECK, PROTO_COMPOSED,
                   PROTO_MULTI_PARTIAL, PROTO_MULTI_FULL)
from .api import (Multilinear, ComposedMultilinear, Sumcheck, SumcheckProof, ComposedSumcheck, ComposedSumcheckProof,
                  MultiComposedSumcheckProver, MultiComposedSumcheckVerifier, MultiComposedProof, SubClaim,
                  FiatShamirTranscript, SparseUnivariatePolynomial, default_context, set_default_context)
from .gkr import Gate, GateType, CircuitLayer, Circuit, GKRProof, GKRProtocol, GKRInstance, LayeredCircuit, SuccintGKRProof, SuccintGKRProtocol
from .kzg import MultilinearKZG, MultilinearKZGProof, TrustedSetup
from . import pairing, utils

__all__ = [n for n in dir() if not n.startswith("_")]
