"""Committed golden vectors (tests/golden/sumcheck_vectors.json, made by tests/golden/make_golden.py from the Python
model): the C oracle must reproduce them on CPU, the CUDA path on the GPU.  The vectors are DERIVED, not emitted by
the Rust reference (it cannot be built here): they pin today's agreed bytes against drift."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import cref
from oracle import pymodel as pm

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "sumcheck_vectors.json")))["vectors"]
PROTO = {"sumcheck": 0, "composed": 1, "multi_partial": 2, "multi_full": 3}


def tables_of(v):
    """-> list of tables (python ints, canonical), in product-major / factor-minor order"""
    if "seed" in v:
        return [pm.synth_table(v["seed"], k, v["n_vars"]).evaluations for k in range(sum(v["degrees"]))]
    return [[int(x, 16) for x in t] for t in v["tables"]]


def check(v, proof_bytes, challenges):
    assert len(proof_bytes) == v["proof_len"]
    assert hashlib.sha256(proof_bytes).hexdigest() == v["proof_sha256"]
    assert [hex(c) for c in challenges] == v["challenges"]
    if v["rounds"]:
        assert proof_bytes[:len(v["round0_hex"]) // 2].hex() == v["round0_hex"]
        assert proof_bytes.hex().endswith(v["last_round_hex"])


@pytest.mark.parametrize("v", VEC, ids=[v["name"] for v in VEC])
def test_c_oracle_reproduces_golden(v, built):
    tabs = tables_of(v)
    flat = cref.ints_to_canon([x for t in tabs for x in t])
    s = int(v["sum"], 16)
    if v["protocol"] != "composed" and "wrong_sum" not in v["name"]:
        assert cref.poly_sum(v["n_vars"], v["degrees"], flat) == s
    proof, ch = cref.prove(PROTO[v["protocol"]], v["n_vars"], v["degrees"], flat, s)
    check(v, proof, ch)


@pytest.mark.gpu
@pytest.mark.parametrize("v", VEC, ids=[v["name"] for v in VEC])
def test_cuda_path_reproduces_golden(v, ctx):
    import zk_cryptography_b200 as zk
    from zk_cryptography_b200 import _lib
    tabs = [zk.to_mont(t) for t in tables_of(v)]
    proto = PROTO[v["protocol"]]
    t = zk.Tables.upload(ctx, v["n_vars"], v["degrees"], tabs)
    try:
        msgs, lens, chal = t.prove(proto, zk.to_mont([int(v["sum"], 16)]))
        check(v, _lib.proof_to_bytes(proto, msgs[0], lens[0]), zk.from_mont(chal[0]) if v["n_vars"] else [])
    finally:
        t.free()
