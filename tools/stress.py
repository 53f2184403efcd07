#!/usr/bin/env python
"""Stress run of the prover: many proofs of random shapes back to back on one context (the resident kernel is started, fed, left and
stopped thousands of times), every proof made twice and compared, a sample compared with the C oracle.  Catches ordering bugs that a
single pass of the test-suite can miss.  Under torch.distributed.run it runs sharded (one rank per GPU); with --devices 0,1,.. it uses
the single-process multi-GPU context.   usage: python tools/stress.py [--seconds 60] [--devices 0,1]"""
import argparse, os, random, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_cryptography_b200 as zk
from zk_cryptography_b200._lib import proof_to_bytes
from oracle import cref

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60)
ap.add_argument("--devices", default=None)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if args.devices:
    ctx = zk.Context(devices=[int(x) for x in args.devices.split(",")])
    lg = ctx.devices().bit_length() - 1
else:
    ctx = zk.Context(local)
    lg = world.bit_length() - 1
    if world > 1:
        import torch, torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(zk.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
rng = random.Random(1234)           # the same sequence on every rank
t_end, n_proofs, n_oracle = time.time() + args.seconds, 0, 0
shapes = [[1], [2], [3], [2, 2], [2, 3], [4], [5], [3, 1], [6]]
stop = False
while not stop:
    n = rng.randint(max(1, lg), 19)
    degs = rng.choice(shapes)
    B = rng.choice([1, 1, 1, 3]) if (world == 1 and not args.devices) else 1
    seed = rng.randrange(1 << 30)
    proto = zk.PROTO_MULTI_PARTIAL
    t = zk.Tables.synth(ctx, n, degs, seed, n_proofs=B)
    s = t.poly_sum()
    a = t.prove(proto, s)
    if rng.random() < 0.3:               # leave a proof half-way through the low-level API, then start over
        t.reset()
        for _ in range(rng.randint(1, n)):
            t.round_evals()
            t.bind(zk.to_mont([rng.randrange(zk.R_MOD) for _ in range(B)]))
    t.reset()
    b = t.prove(proto, s)
    for x, y in zip(a, b):
        assert np.array_equal(x, y), "two proofs of the same tables differ (n=%d degs=%s B=%d seed=%d)" % (n, degs, B, seed)
    if n <= 15 and rng.random() < 0.25 and rank == 0:
        cref.set_threads(cref.max_threads())
        for p in range(B):
            tabs = np.concatenate([cref.synth_table(seed + p, k, n) for k in range(sum(degs))])
            osum = cref.poly_sum(n, degs, tabs)
            assert zk.from_mont(s[p]) == osum
            assert (proof_to_bytes(proto, a[0][p], a[1][p]), zk.from_mont(a[2][p])) == cref.prove(2, n, degs, tabs, osum), "oracle mismatch n=%d degs=%s" % (n, degs)
        n_oracle += B
    t.free()
    n_proofs += 2 * B
    stop = time.time() > t_end
    if world > 1:                        # all ranks must agree on when to stop
        flag = torch.tensor([1 if stop else 0], device="cuda")
        dist.broadcast(flag, 0)
        stop = bool(flag.item())
if rank == 0:
    print("STRESS OK: %d proofs, %d of them against the oracle, world=%d devices=%s" % (n_proofs, n_oracle, world, args.devices))
ctx.close()
