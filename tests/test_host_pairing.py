"""The host-side BLS12-381 pairing (zk_cryptography_b200/pairing.py) and the KZG verifier built on it, on the CPU (no GPU work
is involved in verifying): curve constants, bilinearity, and the reference's own KZG tests -- `verify == true`, and `== false`
with a tampered trusted setup (kzg/src/multilinear_kzg.rs:132-199).  Commitments and openings come from the oracle here
(tests/test_gpu_kzg.py feeds the same verifier with what the GPU computed)."""
import numpy as np
import pytest

from oracle import kzgmodel as k
from zk_cryptography_b200 import pairing as pr
from zk_cryptography_b200.kzg import MultilinearKZG, MultilinearKZGProof, TrustedSetup

R = pr.R


def srs_for(model):
    """the product's TrustedSetup from the oracle's: G1 powers in ark-ec memory form, G2 powers tau_i g2 (trusted_setup.rs:37-46)"""
    return TrustedSetup(np.stack([k.to_ark(p) for p in model.powers_of_tau_in_g1]), [pr.g2_mul(t, pr.G2) for t in model.tau])


def test_generators_and_bilinearity():
    assert pr.g1_on_curve(pr.G1) and pr.g2_on_curve(pr.G2)
    assert pr.g1_mul(R - 1, pr.G1) == pr.g1_neg(pr.G1) and pr.g2_mul(R - 1, pr.G2) == pr.g2_neg(pr.G2)       # both have order r
    assert pr.G1 == k.G1
    e = pr.pairing(pr.G1, pr.G2)
    assert e != pr.F12_ONE and pr.f12_pow(e, R) == pr.F12_ONE                                                # non-degenerate, order r
    a, b = 0x1234567890ABCDEF, 0xFEDCBA0987654321
    assert pr.pairing(pr.g1_mul(a, pr.G1), pr.g2_mul(b, pr.G2)) == pr.f12_pow(e, a * b % R)                   # bilinear
    assert pr.multi_pairing([(pr.g1_mul(a, pr.G1), pr.G2), (pr.g1_neg(pr.G1), pr.g2_mul(a, pr.G2))]) == pr.F12_ONE
    assert pr.pairing(None, pr.G2) == pr.F12_ONE


def test_native_pairing_is_the_python_pairing(built):
    """csrc/host_pairing.hpp through the C ABI (zksc_pairing, zksc_pairing_check; host code, no GPU) against pairing.py: the same
    element of Fq12 coefficient by coefficient, the same decisions, identities, rejected inputs"""
    a, b = 0x1234567890ABCDEF, 0x0FEDCBA987654321
    pa, qb = pr.g1_mul(a, pr.G1), pr.g2_mul(b, pr.G2)
    assert pr.native_pairing(pa, qb) == pr.pairing(pa, qb)
    assert pr.native_pairing(pr.G1, pr.G2) == pr.pairing(pr.G1, pr.G2)
    assert pr.native_pairing_check([(pa, qb), (pr.g1_neg(pr.g1_mul(a * b % R, pr.G1)), pr.G2)]) is True
    assert pr.native_pairing_check([(pa, qb), (pr.g1_neg(pr.g1_mul((a * b + 1) % R, pr.G1)), pr.G2)]) is False
    assert pr.native_pairing_check([(None, pr.G2), (pr.G1, None)]) is True and pr.native_pairing_check([]) is True
    assert pr.native_pairing_check([(pr.G1, pr.G2)]) is False
    from zk_cryptography_b200 import ZkscError
    for bad in ([((1, 2), pr.G2)], [(pr.G1, ((1, 0), (2, 0)))], [((pr.P, 0), pr.G2)]):
        with pytest.raises(ZkscError):
            pr.native_pairing_check(bad)


@pytest.mark.parametrize("prover,tampered,verifier,ev", [
    ([2, 3, 4], [2, 13, 4], [5, 9, 6], [0, 7, 0, 5, 0, 7, 4, 9]),                                                          # test_kzg_1
    ([12, 9, 28, 40], [12, 19, 28, 40], [54, 90, 76, 160], [0, 0, 0, 2, 0, 0, 10, 12, 0, -12, 4, -6, 0, -12, 14, 4]),      # test_kzg_2
])
def test_reference_kzg_verify(prover, tampered, verifier, ev):
    ev = [v % R for v in ev]
    model = k.TrustedSetup(prover)
    commit = k.commitment(ev, model)
    v, proofs = k.open_(ev, verifier, model)
    proof = MultilinearKZGProof(v, np.stack([k.to_ark(p) for p in proofs]))
    assert MultilinearKZG.verify(commit, verifier, proof, srs_for(model)) is True
    assert MultilinearKZG.verify(commit, verifier, proof, srs_for(model), native=False) is True               # the same through pairing.py's integers
    assert MultilinearKZG.verify(commit, verifier, proof, srs_for(k.TrustedSetup(tampered))) is False         # tampered_tau_verify_status == false
    assert MultilinearKZG.verify(commit, verifier, proof, srs_for(k.TrustedSetup(tampered)), native=False) is False
    # ... and with the commitment in ark-ec's memory form, as the device returns it: the whole check inside the library (zksc_kzg_verify)
    ca = k.to_ark(commit)
    assert MultilinearKZG.verify(ca, verifier, proof, srs_for(model)) is True
    assert MultilinearKZG.verify(ca, verifier, proof, srs_for(k.TrustedSetup(tampered))) is False
    assert MultilinearKZG.verify(ca, verifier, MultilinearKZGProof((v + 1) % R, proof.proofs), srs_for(model)) is False
    assert MultilinearKZG.verify(k.to_ark(k.add(commit, k.G1)), verifier, proof, srs_for(model)) is False
    assert MultilinearKZG.verify(ca, [(z + 1) % R for z in verifier], proof, srs_for(model)) is False
    assert MultilinearKZG.verify(commit, verifier, MultilinearKZGProof((v + 1) % R, proof.proofs), srs_for(model)) is False
    assert MultilinearKZG.verify(k.add(commit, k.G1), verifier, proof, srs_for(model)) is False
