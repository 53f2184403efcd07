#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

Every rank builds its shard (index mod G) of the seeded synthetic tables, the G ranks prove together
(per-round exchange of the partial evaluations, residual gather for the last log2 G rounds), and rank 0
compares the proof bytes and challenges with (1) the same proof computed on ONE GPU by an unsharded context
and (2) the CPU oracle -- its prover byte for byte up to 2^24 entries, its verifier plus its own evaluation of the tables
at 2^26 and 2^28 (degree 3: BASELINE config 3).  ZKSC_CHECK_MAX_N caps the sizes.  Prints one line per case and "MULTI-GPU PARITY OK".
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zk_cryptography_b200 as zk  # noqa: E402
from zk_cryptography_b200._lib import proof_to_bytes  # noqa: E402

CASES = [  # (n_vars, degrees, seed, check): "bytes" = the oracle's prover, byte for byte; "verify" = the oracle's verifier + its own
    # streamed evaluation of the tables at the challenges (tests/test_gpu_parity_large.py oracle_verifies); both also compare with one GPU
    (1, [1], 5, "bytes"), (3, [2], 6, "bytes"), (4, [2, 3], 7, "bytes"), (10, [1], 8, "bytes"), (12, [2, 2], 9, "bytes"), (13, [3], 10, "bytes"),
    (14, [5, 1], 11, "bytes"), (18, [2], 12, "bytes"), (20, [3], 13, "bytes"), (22, [2, 2], 14, "bytes"), (24, [2], 15, "bytes"),
    (26, [3], 16, "verify"), (28, [3], 17, "verify"),
]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = zk.Context(local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(zk.Context.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    solo = zk.Context(local) if rank == 0 else None
    if rank == 0:
        print("peer-memory exchange:", ctx.peer_exchange(), flush=True)
    lg = world.bit_length() - 1
    ok = True
    max_n = int(os.environ.get("ZKSC_CHECK_MAX_N", "28"))
    for n, degs, seed, check in CASES:
        if n < lg or n > max_n:
            continue
        t = zk.Tables.synth(ctx, n, degs, seed)
        s = t.poly_sum()
        msgs, lens, chal = t.prove(zk.PROTO_MULTI_PARTIAL, s)
        got = proof_to_bytes(zk.PROTO_MULTI_PARTIAL, msgs[0], lens[0])
        # a second proof on the same handle must give the same bytes (reset path)
        t.reset()
        msgs2, lens2, chal2 = t.prove(zk.PROTO_MULTI_PARTIAL, s)
        same_again = np.array_equal(msgs, msgs2) and np.array_equal(chal, chal2)
        t.free()
        if rank == 0:
            t1 = zk.Tables.synth(solo, n, degs, seed)
            s1 = t1.poly_sum()
            m1, l1, c1 = t1.prove(zk.PROTO_MULTI_PARTIAL, s1)
            want = proof_to_bytes(zk.PROTO_MULTI_PARTIAL, m1[0], l1[0])
            t1.free()
            good = (got == want) and np.array_equal(s, s1) and np.array_equal(chal, c1) and same_again
            from oracle import cref
            cref.set_threads(cref.max_threads())
            if check == "bytes":
                tabs = np.concatenate([cref.synth_table(seed, k, n) for k in range(sum(degs))])
                osum = cref.poly_sum(n, degs, tabs)
                obytes, och = cref.prove(2, n, degs, tabs, osum)
                good = good and zk.from_mont(s[0]) == osum and got == obytes and zk.from_mont(chal[0]) == och
            else:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from test_gpu_parity_large import oracle_verifies
                try:
                    oracle_verifies(n, degs, seed, zk.from_mont(s[0]), got, zk.from_mont(chal[0]))
                except AssertionError as e:
                    print("oracle verification failed:", e, flush=True)
                    good = False
            import hashlib
            print("case n=%d degs=%s G=%d [%s]: %s (%d proof bytes, sha256 %s)" % (n, degs, world, check, "ok" if good else "MISMATCH", len(got),
                                                                                 hashlib.sha256(got).hexdigest()[:16]), flush=True)
            ok = ok and good
    # a batch of independent proofs on the sharded context (per-proof mailboxes, exchange slots and gather stage rows)
    for n, degs, seed, B in ((12, [2], 31, 3), (16, [3, 1], 32, 2)):
        if n < lg:
            continue
        t = zk.Tables.synth(ctx, n, degs, seed, n_proofs=B)
        s = t.poly_sum()
        msgs, lens, chal = t.prove(zk.PROTO_MULTI_PARTIAL, s)
        t.free()
        if rank == 0:
            from oracle import cref
            good = True
            for b in range(B):
                tabs = np.concatenate([cref.synth_table(seed + b, k, n) for k in range(sum(degs))])
                osum = cref.poly_sum(n, degs, tabs)
                good = good and zk.from_mont(s[b]) == osum and (proof_to_bytes(zk.PROTO_MULTI_PARTIAL, msgs[b], lens[b]), zk.from_mont(chal[b])) == cref.prove(2, n, degs, tabs, osum)
            print("batch n=%d degs=%s proofs=%d G=%d [bytes]: %s" % (n, degs, B, world, "ok" if good else "MISMATCH"), flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    if rank == 0:
        print("MULTI-GPU PARITY OK" if ok else "MULTI-GPU PARITY FAILED", flush=True)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
