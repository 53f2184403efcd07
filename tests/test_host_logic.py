"""Host-side logic of the product library (no GPU): field helpers, SHA-256 transcript, sparse polynomial,
transcript replay / verify_partial -- all through the C ABI, checked against the Python model."""
import hashlib
import os
import random

import numpy as np
import pytest

import zk_cryptography_b200 as zk
from zk_cryptography_b200 import _lib
from oracle import pymodel as pm

R = pm.R_MOD
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    return built


def test_field_helpers_roundtrip_and_ops():
    rng = random.Random(1)
    vals = [0, 1, R - 1, R - 2, 2**255 % R] + [rng.randrange(R) for _ in range(200)]
    m = zk.to_mont(vals)
    assert zk.from_mont(m) == vals
    # Montgomery form = v * 2^256 mod r, little-endian limbs (ark-ff memory layout)
    assert _lib.limbs_to_ints(m) == [(v << 256) % R for v in vals]
    L = zk.lib()
    out = np.zeros(4, dtype=np.uint64)
    for a, b in zip(vals, reversed(vals)):
        A, B = zk.to_mont(a), zk.to_mont(b)
        L.zksc_fr_mul(_lib.p64(A), _lib.p64(B), _lib.p64(out)); assert zk.from_mont(out) == a * b % R
        L.zksc_fr_add(_lib.p64(A), _lib.p64(B), _lib.p64(out)); assert zk.from_mont(out) == (a + b) % R
        L.zksc_fr_sub(_lib.p64(A), _lib.p64(B), _lib.p64(out)); assert zk.from_mont(out) == (a - b) % R
        be = np.zeros(32, dtype=np.uint8)
        L.zksc_fr_to_be_bytes(_lib.p64(A), _lib.p8(be)); assert be.tobytes() == pm.be32(a)
    L.zksc_fr_from_u64(100, _lib.p64(out)); assert zk.from_mont(out) == 100
    for raw in (b"\xff" * 32, bytes(range(32)), (R).to_bytes(32, "big"), (R - 1).to_bytes(32, "big")):
        buf = np.frombuffer(raw, dtype=np.uint8).copy()
        L.zksc_fr_from_be_bytes_mod_order(_lib.p8(buf), _lib.p64(out))
        assert zk.from_mont(out) == int.from_bytes(raw, "big") % R


def test_sha256_and_transcript():
    for msg in (b"", b"abc", b"a" * 55, b"a" * 56, b"a" * 63, b"a" * 64, b"a" * 65, bytes(range(256)) * 41):
        t = zk.FiatShamirTranscript()
        t.commit(msg)
        c1 = t.challenge()
        assert c1 == hashlib.sha256(msg).digest()
        assert t.challenge() == hashlib.sha256(c1).digest()      # fiat_shamir.rs:21-25: digest is fed back
    t, p = zk.FiatShamirTranscript(), pm.FiatShamirTranscript()
    for chunk in (b"x" * 7, b"y" * 100, b"", b"z" * 64):          # streaming updates
        t.commit(chunk); p.commit(chunk)
    assert t.evaluate_n_challenge_into_field(3) == p.evaluate_n_challenge_into_field(3)


def test_sparse_interpolate_add_evaluate():
    rng = random.Random(2)
    S = zk.SparseUnivariatePolynomial
    for trial in range(60):
        d = rng.randint(0, 8)
        kind = trial % 3
        ys = [rng.randrange(R) if kind == 0 else rng.randrange(3) if kind == 1 else 0 for _ in range(d + 1)]
        want = pm.SparseUnivariatePolynomial.interpolation(pm.convert_round_poly_to_uni_poly_format(ys))
        got = S.interpolation_evals(ys)
        assert got.monomial == want.monomial
        ys2 = [rng.randrange(R) if kind != 2 else 0 for _ in range(rng.randint(1, 6))]
        w2 = pm.SparseUnivariatePolynomial.interpolation(pm.convert_round_poly_to_uni_poly_format(ys2))
        g2 = S.interpolation_evals(ys2)
        assert (got + g2).monomial == want.add(w2).monomial
        x = rng.randrange(R)
        assert got.evaluate(x) == want.evaluate(x)
        assert got.to_bytes() == want.to_bytes()
    # zero sums are kept by Add; zero coefficients are dropped by interpolation
    a, b = S([(5, 1)]), S([(R - 5, 1)])
    assert (a + b).monomial == [(0, 1)]
    assert S.interpolation_evals([0, 2]).monomial == [(2, 1)]
    assert S.interpolation_evals([5, 7, 13]).monomial == [(5, 0), (2, 2)]


def _to_raw(round_polys, stride):
    n = len(round_polys)
    msgs = np.zeros((n, stride, 4), dtype=np.uint64)
    lens = np.zeros(n, dtype=np.uint32)
    for r, rp in enumerate(round_polys):
        flat = [v for cp in rp for v in cp] if rp and isinstance(rp[0], tuple) else list(rp)
        if flat:
            msgs[r, :len(flat)] = zk.to_mont(flat)
        lens[r] = len(rp)
    return msgs, lens


def test_verify_rounds_replays_oracle_proofs():
    rng = random.Random(3)
    for trial in range(10):
        n = rng.randint(1, 5)
        degs = [rng.randint(1, 3) for _ in range(rng.randint(1, 3))]
        polys = [pm.ComposedMultilinear([pm.Multilinear([rng.randrange(R) for _ in range(1 << n)]) for _ in range(d)]) for d in degs]
        s = pm.MultiComposedSumcheckProver.calculate_poly_sum(polys)
        pr, ch = pm.MultiComposedSumcheckProver.prove_partial(polys, s)
        stride = 2 * (max(degs) + 1)
        msgs, lens = _to_raw([rp.monomial for rp in pr.round_polys], stride)
        sub, chal = _lib.verify_rounds(zk.PROTO_MULTI_PARTIAL, zk.to_mont(s), msgs, lens)
        want = pm.MultiComposedSumcheckVerifier.verify_partial(pr)
        assert zk.from_mont(sub) == want.sum and zk.from_mont(chal) == want.challenges == ch
        assert _lib.proof_to_bytes(zk.PROTO_MULTI_PARTIAL, msgs, lens) == pr.to_bytes()
        with pytest.raises(zk.ZkscError) as e:
            _lib.verify_rounds(zk.PROTO_MULTI_PARTIAL, zk.to_mont((s + 1) % R), msgs, lens)
        assert e.value.code == -7  # Err("Verification failed"), multi_composed_sumcheck.rs:170
        # full `prove`: tables absorbed first
        prf, chf = pm.MultiComposedSumcheckProver.prove(polys, s)
        msgs, lens = _to_raw([rp.monomial for rp in prf.round_polys], stride)
        sub, chal = _lib.verify_rounds(zk.PROTO_MULTI_FULL, zk.to_mont(s), msgs, lens, pm.composed_poly_to_bytes(polys))
        assert zk.from_mont(chal) == chf
    # Sumcheck / ComposedSumcheck message formats
    ev = [rng.randrange(R) for _ in range(16)]
    sc = pm.Sumcheck(pm.Multilinear(ev)); sc.poly_sum()
    pr, ch = sc.prove()
    msgs, lens = _to_raw([u.evaluations for u in pr.univariate_poly], 2)
    sub, chal = _lib.verify_rounds(zk.PROTO_SUMCHECK, zk.to_mont(sc.sum), msgs, lens)
    assert zk.from_mont(chal) == ch and zk.from_mont(sub) == pm.Multilinear(ev).evaluation(ch)
    cs = pm.ComposedSumcheck(pm.ComposedMultilinear([pm.Multilinear(ev), pm.Multilinear(ev[::-1])]))
    pr, ch = cs.prove()
    msgs, lens = _to_raw(pr.round_polys, 3)
    sub, chal = _lib.verify_rounds(zk.PROTO_COMPOSED, zk.to_mont(pm.ComposedSumcheck.calculate_poly_sum(cs.poly)), msgs, lens)
    assert zk.from_mont(chal) == ch and zk.from_mont(sub) == cs.poly.evaluation(ch)


def test_synth_entry_matches_model():
    out = np.zeros(4, dtype=np.uint64)
    for seed, tab, i in ((7, 1, 0), (7, 1, 5), (2026, 3, 123456), (2**64 - 1, 0, 2**40 + 3)):
        zk.lib().zksc_synth_entry(seed, tab, i, _lib.p64(out))
        assert zk.from_mont(out) == pm.synth_entry(seed, tab, i)


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    if zk.lib().zksc_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(zk.ZkscError) as e:
        zk.Context(0)
    assert e.value.code == -1


def test_sha256_scalar_and_shani_paths_agree():
    """The library hashes with the x86 SHA extensions when the CPU has them (host_field.hpp block_shani) and with the portable
    compression otherwise (ZKSC_NO_SHANI=1 forces it; read once per process, hence the subprocess): both must equal hashlib."""
    import subprocess, sys
    code = (
        "import sys, hashlib, ctypes, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import zk_cryptography_b200 as zk\n"
        "L = zk.lib()\n"
        "for n in list(range(0, 200)) + [255, 256, 1000, 4097]:\n"
        "    msg = bytes((7 * i + n) & 255 for i in range(n))\n"
        "    t = L.zksc_transcript_new(); L.zksc_transcript_commit(t, msg, n)\n"
        "    out = np.zeros(32, dtype=np.uint8); L.zksc_transcript_challenge(t, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))\n"
        "    assert out.tobytes() == hashlib.sha256(msg).digest(), n\n"
        "    L.zksc_transcript_free(t)\n"
        "print('ok')\n") % ROOT
    for env_extra in ({}, {"ZKSC_NO_SHANI": "1"}):
        env = dict(os.environ, **env_extra)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-1500:]


def test_round_slots_to_evals():
    """host::round_slots_to_evals (what finish_round applies to every round's device results): degree-2 products arrive as
    (h(0), h(1), h(inf)), degree-3 products as (h(0), h(1), h(-1), h(inf)); the helper must return the evaluations at 0..d --
    checked on random polynomials with Python big integers; other degrees pass through untouched."""
    rng = random.Random(31)
    L = zk.lib()
    for trial in range(200):
        for d in (2, 3):
            c = [rng.randrange(R) for _ in range(d + 1)]
            if trial < 4:
                c = [[0] * (d + 1), [R - 1] * (d + 1), [1] + [0] * d, [0] * d + [1]][trial]
            h = lambda t: sum(ci * pow(t, i, R) for i, ci in enumerate(c)) % R
            slots = [h(0), h(1), c[2]] if d == 2 else [h(0), h(1), h(R - 1), c[3]]
            arr = zk.to_mont(slots)
            L.zksc_round_slots_to_evals(d, _lib.p64(arr))
            assert zk.from_mont(arr) == [h(t) for t in range(d + 1)]
    for d in (1, 4, 5):
        vals = [rng.randrange(R) for _ in range(d + 1)]
        arr = zk.to_mont(vals)
        L.zksc_round_slots_to_evals(d, _lib.p64(arr))
        assert zk.from_mont(arr) == vals
