"""Host-side mirror of the reference's multilinear KZG over BLS12-381 (kzg/src/multilinear_kzg.rs, kzg/src/trusted_setup.rs) for
the part that is table-sized work: `commitment` and `open` run on the GPU through zksc_g1_msm / zksc_kzg_open.  G1 points are held
in ark-ec's in-memory form (numpy (18,) uint64: Jacobian X, Y, Z in Montgomery form), so a reference-side `Vec<G1Projective>` maps
onto a (n, 18) array unchanged.  `verify` is n + 1 pairings of single points -- host-side, constant-size work like the transcript: it
runs in `pairing.py` (plain integers), not on the GPU."""
import numpy as np

from . import pairing
from ._lib import ZkscError, from_mont, lib, p64, to_mont
from .api import Multilinear, default_context


class TrustedSetup:
    """TrustedSetup { powers_of_tau_in_g1, powers_of_tau_in_g2 } (kzg/src/trusted_setup.rs:10-13): supplied by the caller -- generating
    it is a one-off the reference does with 2^n scalar multiplications of the generator (:24-35); tests build it with the oracle.
    powers_of_tau_in_g2 (only the verifier needs it): one affine G2 point per variable, ((x0, x1), (y0, y1)) with Fq2 = x0 + x1 u."""

    def __init__(self, powers_of_tau_in_g1, powers_of_tau_in_g2=None):
        self.powers_of_tau_in_g1 = np.ascontiguousarray(powers_of_tau_in_g1, dtype=np.uint64).reshape(-1, 18)
        self.powers_of_tau_in_g2 = list(powers_of_tau_in_g2) if powers_of_tau_in_g2 is not None else None


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

    @staticmethod
    def verify(commit, verifier_points, proof: MultilinearKZGProof, srs: TrustedSetup, native=True):  # multilinear_kzg.rs:90-116
        """e(commit - evaluation g1, g2) == sum_i e(proof_i, tau_i g2 - z_i g2)   (sum_pairing_results, kzg/src/utils.rs:42-61), checked
        as one product of n + 1 Miller loops and a single final exponentiation -- by the library's host code (zksc_pairing_check,
        csrc/host_pairing.hpp) or, native=False, by the Python integers of pairing.py (the cross-check of the tests)."""
        if srs.powers_of_tau_in_g2 is None:
            raise ZkscError(-3, "the trusted setup carries no G2 powers: cannot verify")
        proofs = np.ascontiguousarray(proof.proofs, dtype=np.uint64).reshape(-1, 18)
        if not (len(verifier_points) == len(srs.powers_of_tau_in_g2) == proofs.shape[0]):      # assert_eq! "Length mismatch", utils.rs:49-50
            raise ZkscError(-3, "Length mismatch")
        if native and isinstance(commit, np.ndarray):
            # everything in the library's host code: zksc_kzg_verify (group arithmetic, Miller loops, final exponentiation)
            import ctypes
            ev = np.ascontiguousarray(proof.evaluation, dtype=np.uint64) if isinstance(proof.evaluation, np.ndarray) else to_mont([int(proof.evaluation)])
            g2 = np.array([w for q in srs.powers_of_tau_in_g2 for w in pairing._pack_g2(q)], dtype=np.uint64)
            ok = ctypes.c_int(0)
            rc = lib().zksc_kzg_verify(p64(np.ascontiguousarray(commit, dtype=np.uint64)), p64(to_mont([int(z) for z in verifier_points])), p64(ev), p64(proofs), p64(g2),
                                       proofs.shape[0], ctypes.byref(ok))
            if rc != 0:
                raise ZkscError(rc, "zksc_kzg_verify: malformed point")
            return bool(ok.value)
        c = pairing.g1_from_ark(commit) if isinstance(commit, np.ndarray) else commit
        v = int(from_mont(np.ascontiguousarray(proof.evaluation, dtype=np.uint64))) if isinstance(proof.evaluation, np.ndarray) else int(proof.evaluation)
        lhs_point = pairing.g1_add(c, pairing.g1_neg(pairing.g1_mul(v, pairing.G1)))
        pairs = [(lhs_point, pairing.g2_neg(pairing.G2))]
        for i, z in enumerate(verifier_points):
            pairs.append((pairing.g1_from_ark(proofs[i]), pairing.g2_add(srs.powers_of_tau_in_g2[i], pairing.g2_neg(pairing.g2_mul(int(z), pairing.G2)))))
        return pairing.native_pairing_check(pairs) if native else pairing.multi_pairing(pairs) == pairing.F12_ONE
