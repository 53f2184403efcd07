"""Byte parity of the CUDA path against the C oracle AT BASELINE.json's configuration sizes, and oracle-side verification of
the proofs that are too large for the oracle's prover.

Every case here drives the paths the benchmark numbers run through -- several grid-stride iterations per thread, slot refill of
the TMA-staged kernels, the resident rounds kernel from its first round to the last -- which the small cases of
tests/test_gpu_parity.py do not (592 CTAs x 128 threads cover 2^16 pairs in one iteration).

  * n <= 24: proof bytes and challenges must equal oracle/zkref.c's (`cref.prove`), i.e. the restatement of
    sumcheck/src/sumcheck.rs:29-61 and sumcheck/src/composed/multi_composed_sumcheck.rs:64-120.
  * n = 26, 28 (degree 3; 2^28 is BASELINE config 3 and the north_star target): the oracle replays the transcript
    (`cref.verify_partial`, multi_composed_sumcheck.rs:151-181: challenges re-derived from the proof bytes, p(0) + p(1) chain) and
    closes the final claim against ITS OWN evaluation of the seeded tables at the challenges (`cref.eval_synth`, a streamed
    Multilinear::evaluation, evaluation_form.rs:162-175).  A round polynomial that differed from the honest prover's would
    survive that only with probability ~ n d / |F|, so an accepted proof is the reference prover's proof.
File:line references are to /root/reference."""
import os

import numpy as np
import pytest

import zk_cryptography_b200 as zk
from zk_cryptography_b200 import _lib
from oracle import cref
from oracle import pymodel as pm

pytestmark = pytest.mark.gpu
R = pm.R_MOD


def _oracle(proto, n, degs, seed):
    cref.set_threads(cref.max_threads())
    tabs = np.concatenate([cref.synth_table(seed, k, n) for k in range(sum(degs))])
    s = cref.poly_sum(n, degs, tabs)
    out = cref.prove(proto, n, degs, tabs, s)
    cref.set_threads(1)
    return s, out


def _gpu(ctx, proto, n, degs, seed, n_proofs=1):
    t = zk.Tables.synth(ctx, n, degs, seed, n_proofs=n_proofs)
    try:
        sums = t.poly_sum()
        msgs, lens, chal = t.prove(proto, sums)
        return [(zk.from_mont(sums[b]), _lib.proof_to_bytes(proto, msgs[b], lens[b]), zk.from_mont(chal[b])) for b in range(n_proofs)]
    finally:
        t.free()


@pytest.mark.parametrize("n,degs,proto,label", [
    (20, [1], zk.PROTO_SUMCHECK, "c1: Sumcheck::prove, 2^20"),
    (24, [2], zk.PROTO_MULTI_PARTIAL, "c2: prove_partial, degree 2, 2^24"),
    (24, [3], zk.PROTO_MULTI_PARTIAL, "degree 3 at 2^24 (c3's kernels, several grid-stride iterations)"),
    (21, [2, 2], zk.PROTO_MULTI_PARTIAL, "GKR layer shape: two degree-2 products sharing launches"),
    (20, [2, 3], zk.PROTO_MULTI_PARTIAL, "the reference bench's shape: a degree-2 plus a degree-3 product"),
])
def test_config_size_byte_parity(ctx, n, degs, proto, label):
    s, (want_bytes, want_chal) = _oracle(proto, n, degs, 7000 + n)
    got = _gpu(ctx, proto, n, degs, 7000 + n)[0]
    assert got[0] == s, label + ": calculate_poly_sum differs"
    assert got[2] == want_chal, label + ": challenges differ"
    assert got[1] == want_bytes, label + ": proof bytes differ"


def test_c5_shape_batch_byte_parity(ctx):
    """BASELINE config 5's shape (independent 2^22-entry degree-2 proofs in batched launches), 4 proofs, every one against the oracle"""
    n, degs, B, seed = 22, [2], 4, 8100
    got = _gpu(ctx, zk.PROTO_MULTI_PARTIAL, n, degs, seed, n_proofs=B)
    for b in range(B):
        s, (want_bytes, want_chal) = _oracle(2, n, degs, seed + b)
        assert got[b] == (s, want_bytes, want_chal), "proof %d of the batch differs from the oracle" % b


@pytest.mark.parametrize("env", [{"ZKSC_STAGED_FOLD": "1"}, {"ZKSC_NO_TAIL": "1"}, {"ZKSC_NO_STAGED": "1"}, {"ZKSC_TAIL_WORK": "1"},
                                 {"ZKSC_TAIL_WORK": "1000000000000"}, {"ZKSC_ROUND_STATIC": "1"}, {"ZKSC_RES_STATIC": "1"},
                                 {"ZKSC_ROUND_STATIC": "1", "ZKSC_RES_STATIC": "1", "ZKSC_NO_STAGED": "1"}, {"ZKSC_STAGED_FOLD": "1", "ZKSC_NO_TAIL": "1"}])
@pytest.mark.parametrize("n,degs", [(22, [2]), (22, [3])])
def test_kernel_variants_byte_parity(built, env, n, degs):
    """the alternative data paths (TMA-staged fold rounds, no resident kernel, no staged kernels, resident kernel only for the
    last rounds / from round 1 on, fixed split instead of chunks handed out from counters) must give the oracle's bytes too; a fresh
    context reads the switches"""
    s, (want_bytes, want_chal) = _oracle(2, n, degs, 7700 + n)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        c = zk.Context(0)
        try:
            got = _gpu(c, zk.PROTO_MULTI_PARTIAL, n, degs, 7700 + n)[0]
        finally:
            c.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert got == (s, want_bytes, want_chal), "variant %r differs from the oracle" % (env,)


@pytest.mark.parametrize("env", [{"ZKSC_DYN_MAX_GROUPS": "64"}, {"ZKSC_DYN_MAX_GROUPS": "0"}])
def test_batch_with_and_without_work_hand_out(built, env):
    """a batch of 6 proofs of two products each (12 groups per launch): chunks from counters in every launch, or in none"""
    n, degs, B, seed = 18, [2, 2], 6, 8300
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        c = zk.Context(0)
        try:
            got = _gpu(c, zk.PROTO_MULTI_PARTIAL, n, degs, seed, n_proofs=B)
        finally:
            c.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for b in range(B):
        s, (want_bytes, want_chal) = _oracle(2, n, degs, seed + b)
        assert got[b] == (s, want_bytes, want_chal), "proof %d of the batch differs from the oracle (%r)" % (b, env)


def oracle_verifies(n, degs, seed, s, proof_bytes, chal):
    """The oracle's verifier + the oracle's own final evaluation (one product of degs[0] tables): oracle/cref.py verify_synth_proof."""
    assert len(degs) == 1
    cref.set_threads(cref.max_threads())
    ok, why = cref.verify_synth_proof(n, degs[0], seed, s, proof_bytes, chal)
    cref.set_threads(1)
    assert ok, why
    return True


@pytest.mark.parametrize("n", [26, 28])
def test_target_size_degree3_oracle_verified(ctx, n):
    """BASELINE config 3 / the north_star target (2^28 entries, degree 3) and 2^26: see the module docstring"""
    degs, seed = [3], 9100 + n
    try:
        s, proof, chal = _gpu(ctx, zk.PROTO_MULTI_PARTIAL, n, degs, seed)[0]
    except zk.ZkscError as e:
        if e.code == -6:
            pytest.skip("not enough device memory for 2^%d x 3 tables" % n)
        raise
    assert oracle_verifies(n, degs, seed, s, proof, chal)


def test_oracle_verification_rejects_a_wrong_proof(ctx):
    """the full-size check above has teeth: a proof of other tables, or a flipped byte, is rejected by it"""
    n, degs, seed = 16, [3], 555
    s, proof, chal = _gpu(ctx, zk.PROTO_MULTI_PARTIAL, n, degs, seed)[0]
    assert oracle_verifies(n, degs, seed, s, proof, chal)
    with pytest.raises(AssertionError):
        oracle_verifies(n, degs, seed + 1, s, proof, chal)
    bad = bytearray(proof)
    bad[100] ^= 1
    with pytest.raises(AssertionError):
        oracle_verifies(n, degs, seed, s, bytes(bad), chal)
