"""Debug: per-round evaluations of a sharded context (peer exchange) against a sharded context forced onto the
NCCL path and an unsharded one."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zk_cryptography_b200 as zk

def mk_ctx(local, world, rank):
    ctx = zk.Context(local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(zk.Context.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    return ctx

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    a = mk_ctx(local, world, rank)
    os.environ["ZKSC_NO_P2P"] = "1"
    b = mk_ctx(local, world, rank)
    solo = zk.Context(local)
    print(rank, "p2p a/b", a.peer_exchange(), b.peer_exchange(), flush=True)
    for n, degs in ((3, [2]), (6, [2, 3]), (12, [3])):
        ta, tb, ts = (zk.Tables.synth(c, n, degs, 77) for c in (a, b, solo))
        for rnd in range(n):
            ea, eb, es = ta.round_evals(), tb.round_evals(), ts.round_evals()
            okab, okas = np.array_equal(ea, eb), np.array_equal(ea, es)
            print("rank %d n=%d degs=%s round %d: p2p==nccl %s  p2p==solo %s nccl==solo %s" % (rank, n, degs, rnd, okab, okas, np.array_equal(eb, es)), flush=True)
            if not okab:
                print(rank, "p2p ", [hex(v)[:14] for v in zk.from_mont(ea[0])], flush=True)
                print(rank, "nccl", [hex(v)[:14] for v in zk.from_mont(eb[0])], flush=True)
            ch = zk.to_mont([1234567 + rnd])
            ta.bind(ch); tb.bind(ch); ts.bind(ch)
        dist.barrier()
    dist.destroy_process_group()
main()
