"""Host-side multi-GPU logic on CPU, world_size 2 (gloo): the sharding convention of DESIGN.md
("rank g holds the entries i with i mod G == g"), the per-round exchange (gather every rank's partial round
evaluations, add mod r identically on every rank) and the residual gather must reproduce the unsharded
proof bytes.  The per-shard arithmetic comes from the Python oracle here (test infrastructure); on GPUs the
same control flow runs in zksc.cu (round_evals_impl / gather_tail) with ncclAllGather."""
import os
import random
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ints_to_tensor(vals):
    out = torch.zeros(len(vals), 4, dtype=torch.int64)
    for i, v in enumerate(vals):
        for k in range(4):
            w = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
            out[i, k] = w - (1 << 64) if w >= (1 << 63) else w
    return out


def _tensor_to_ints(t):
    return [sum(((int(t[i, k]) + (1 << 64)) % (1 << 64)) << (64 * k) for k in range(4)) for i in range(t.shape[0])]


def _worker(rank, world, port, n, degs, seed, q):
    sys.path.insert(0, ROOT)
    from oracle import pymodel as pm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    R = pm.R_MOD
    rng = random.Random(seed)
    full = [[[rng.randrange(R) for _ in range(1 << n)] for _ in range(d)] for d in degs]
    # shard: local table T_g[i'] = T[i' * G + g]
    local = [pm.ComposedMultilinear([pm.Multilinear(t[rank::world]) for t in tp]) for tp in full]
    polys_full = [pm.ComposedMultilinear([pm.Multilinear(t) for t in tp]) for tp in full]
    s = pm.MultiComposedSumcheckProver.calculate_poly_sum(polys_full)
    tr = pm.FiatShamirTranscript()
    tr.commit(pm.be32(s))
    proof_bytes, challenges = b"", []
    cur, gathered = local, False
    for rnd in range(n):
        if not gathered and cur[0].n_vars() == 0:
            # every rank is down to one entry per table: gather (entry index == rank) and continue replicated
            mine = _ints_to_tensor([m.evaluations[0] for p in cur for m in p.polys])
            parts = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            vals = [_tensor_to_ints(p) for p in parts]
            k, nxt = 0, []
            for p in cur:
                ms = []
                for _ in p.polys:
                    ms.append(pm.Multilinear([vals[g][k] for g in range(world)]))
                    k += 1
                nxt.append(pm.ComposedMultilinear(ms))
            cur, gathered = nxt, True
        evals = [v for p in cur for v in pm.round_evals(p)]
        if not gathered:
            mine = _ints_to_tensor(evals)
            parts = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)                       # the only per-round exchange
            allv = [_tensor_to_ints(p) for p in parts]
            evals = [sum(allv[g][i] for g in range(world)) % R for i in range(len(evals))]
        round_poly, off = pm.SparseUnivariatePolynomial.zero(), 0
        for p in cur:
            d = p.max_degree()
            round_poly = round_poly.add(pm.SparseUnivariatePolynomial.interpolation(pm.convert_round_poly_to_uni_poly_format(evals[off:off + d + 1])))
            off += d + 1
        tr.commit(round_poly.to_bytes())
        proof_bytes += round_poly.to_bytes()
        r = tr.evaluate_challenge_into_field()
        challenges.append(r)
        cur = [p.partial_evaluation(r, 0) for p in cur]
    want, want_ch = pm.MultiComposedSumcheckProver.prove_partial(polys_full, s)
    q.put((rank, proof_bytes == want.to_bytes(), challenges == want_ch))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,degs", [(1, [2]), (3, [1]), (5, [2, 3]), (6, [3])])
def test_sharded_prover_matches_unsharded_world2(n, degs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randint(0, 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, degs, 77, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res
