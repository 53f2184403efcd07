#!/usr/bin/env python
"""Generates tests/golden/sumcheck_vectors.json from the independent Python model (oracle/pymodel.py).

    python tests/golden/make_golden.py

STATUS: these vectors are DERIVED (Python big-int restatement of the reference), not emitted by the Rust binary:
the reference cannot be built in this image (no cargo/rustc; ark-ff / ark-test-curves / sha2 are not vendored).
They freeze today's agreed behaviour of both oracles so that drift in the C oracle or the CUDA path is caught.
Inputs come from the seeded generator (pymodel.synth_entry) or are the reference's own test inputs (cited)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pymodel as pm  # noqa: E402

ML = pm.Multilinear
CM = pm.ComposedMultilinear


def synth_polys(seed, n, degs):
    polys, k = [], 0
    for d in degs:
        polys.append(CM([pm.synth_table(seed, k + j, n) for j in range(d)]))
        k += d
    return polys


def record(name, proto, polys, note, wrong_sum=False):
    if proto == "sumcheck":
        sc = pm.Sumcheck(polys[0].polys[0])
        sc.poly_sum()
        proof, ch = sc.prove()
        s = sc.sum
        rounds = [pm.vec_to_bytes(u.evaluations) for u in proof.univariate_poly]
    elif proto == "composed":
        cs = pm.ComposedSumcheck(polys[0])
        s = pm.ComposedSumcheck.calculate_poly_sum(polys[0])
        proof, ch = cs.prove()
        rounds = [pm.vec_to_bytes(r) for r in proof.round_polys]
    else:
        s = pm.MultiComposedSumcheckProver.calculate_poly_sum(polys)
        claimed = (s + 7) % pm.R_MOD if wrong_sum else s
        fn = pm.MultiComposedSumcheckProver.prove if proto == "multi_full" else pm.MultiComposedSumcheckProver.prove_partial
        proof, ch = fn(polys, claimed)
        rounds = [rp.to_bytes() for rp in proof.round_polys]
        s = claimed
    blob = b"".join(rounds)
    tables = [[hex(v) for v in m.evaluations] for p in polys for m in p.polys]
    return {"name": name, "protocol": proto, "note": note, "sum": hex(s), "n_vars": polys[0].n_vars(), "tables": tables,
            "degrees": [p.max_degree() for p in polys], "rounds": len(rounds), "proof_len": len(blob),
            "proof_sha256": hashlib.sha256(blob).hexdigest(), "round0_hex": rounds[0].hex() if rounds else "",
            "last_round_hex": rounds[-1].hex() if rounds else "", "challenges": [hex(c) for c in ch]}


def main():
    out = []
    # reference's own inputs
    out.append(record("ref_sumcheck_8", "sumcheck", [CM([ML([0, 0, 2, 7, 3, 3, 6, 11])])], "sumcheck/src/sumcheck.rs:127-136"))
    out.append(record("ref_bench_0_255", "sumcheck", [CM([ML(list(range(256)))])], "sumcheck/benches/sumcheck_benchmark.rs:7-12 (values 0..=255)"))
    f = ML([0, 0, 0, 2, 0, 10, 0, 17])
    g = ML([0, 1, 2, 3, 4, 5, 6, 7])
    out.append(record("composed_d2_8", "composed", [CM([f, g])], "3-variable product (shape of composed_sumcheck.rs:143-164)"))
    # seeded synthetic cases (entry = pymodel.synth_entry(seed, table, i)); seed and shape are the inputs
    for seed, n, degs, proto in [(11, 6, [1], "sumcheck"), (12, 7, [3], "composed"), (13, 8, [2], "multi_partial"), (14, 7, [2, 3], "multi_partial"),
                                 (15, 6, [2, 2], "multi_full"), (16, 5, [1, 4, 2], "multi_partial"), (17, 10, [3], "multi_partial"), (18, 9, [5], "composed")]:
        r = record("synth_s%d_n%d_%s_%s" % (seed, n, "x".join(map(str, degs)), proto), proto, synth_polys(seed, n, degs), "seeded synthetic tables")
        r["seed"] = seed
        del r["tables"]              # regenerated from the seed
        out.append(r)
    r = record("synth_s19_n6_2x3_wrong_sum", "multi_partial", synth_polys(19, 6, [2, 3]), "caller-supplied sum is NOT the true sum (absorbed as given, multi_composed_sumcheck.rs:70)", wrong_sum=True)
    r["seed"] = 19
    del r["tables"]
    out.append(r)
    # structured tables: exercise zero-dropping / zero-sum-keeping of the sparse round polynomial
    z = ML([0] * 16)
    one = ML([1] * 16)
    bits = ML([(i * 7 + 3) % 2 for i in range(16)])
    out.append(record("all_zero_d2", "multi_partial", [CM([z, z])], "all-zero tables: every round polynomial is empty"))
    out.append(record("ones_times_bits_plus_zero", "multi_partial", [CM([one, bits]), CM([z, one, bits])], "0/1 tables, a zero product: zero coefficients dropped per product, zero sums kept on add"))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sumcheck_vectors.json"), "w") as fjs:
        json.dump({"generator": "tests/golden/make_golden.py (oracle/pymodel.py)", "status": "derived, not emitted by the Rust reference (parity unpinned at protocol level)",
                   "field": "BLS12-381 Fr", "vectors": out}, fjs, indent=1)
    print("wrote", len(out), "vectors")


if __name__ == "__main__":
    main()
