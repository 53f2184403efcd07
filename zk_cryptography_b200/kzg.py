"""Host-side mirror of the reference's multilinear KZG over BLS12-381 (kzg/src/multilinear_kzg.rs, kzg/src/trusted_setup.rs) for
the part that is table-sized work: `commitment` and `open` run on the GPU through zksc_g1_msm / zksc_kzg_open.  G1 points are held
in ark-ec's in-memory form (numpy (18,) uint64: Jacobian X, Y, Z in Montgomery form), so a reference-side `Vec<G1Projective>` maps
onto a (n, 18) array unchanged.  The pairing check of `verify` is the caller's (ark-ec): not part of this path."""
import numpy as np

from ._lib import ZkscError, lib, p64, to_mont
from .api import Multilinear, default_context


class TrustedSetup:
    """TrustedSetup { powers_of_tau_in_g1 } (kzg/src/trusted_setup.rs:10-13): supplied by the caller -- generating it is a
    one-off the reference does with 2^n scalar multiplications of the generator (:24-35); tests build it with the oracle."""

    def __init__(self, powers_of_tau_in_g1):
        self.powers_of_tau_in_g1 = np.ascontiguousarray(powers_of_tau_in_g1, dtype=np.uint64).reshape(-1, 18)


class MultilinearKZGProof:  # multilinear_kzg.rs:17-21
    def __init__(self, evaluation, proofs):
        self.evaluation, self.proofs = evaluation, proofs


class MultilinearKZG:
    @staticmethod
    def commitment(poly: Multilinear, srs: TrustedSetup, ctx=None):  # multilinear_kzg.rs:33-48
        ctx = ctx or default_context()
        n = poly.evaluations.shape[0]
        if srs.powers_of_tau_in_g1.shape[0] != n:      # assert_eq! :36-41
            raise ZkscError(-3, "The length of powers_of_tau_in_g1 and the length of the evaluations of the polynomial should tally!")
        out = np.zeros(18, dtype=np.uint64)
        ctx.check(lib().zksc_g1_msm(ctx._h, p64(poly.evaluations), p64(srs.powers_of_tau_in_g1), n, p64(out)))
        return out

    @staticmethod
    def open(poly: Multilinear, evaluation_points, srs: TrustedSetup, ctx=None):  # multilinear_kzg.rs:50-88
        ctx = ctx or default_context()
        n = poly.evaluations.shape[0]
        if len(evaluation_points) != poly.n_vars or srs.powers_of_tau_in_g1.shape[0] != n:
            raise ZkscError(-3, "one evaluation point per variable, one power of tau per evaluation")
        pts = to_mont([int(v) for v in evaluation_points]).reshape(-1, 4)
        ev = np.zeros(4, dtype=np.uint64)
        proofs = np.zeros((poly.n_vars, 18), dtype=np.uint64)
        ctx.check(lib().zksc_kzg_open(ctx._h, p64(poly.evaluations), poly.n_vars, p64(pts), p64(srs.powers_of_tau_in_g1), p64(ev), p64(proofs)))
        return MultilinearKZGProof(ev, proofs)
