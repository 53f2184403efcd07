"""Sharded proving on >= 2 GPUs against one GPU and the oracle, under pytest: spawns tools/multi_gpu_check.py through
torch.distributed.run on every power-of-two GPU count the box has (skipped on a single-GPU box).  The per-round exchange of
the partial evaluations over NVLink peer memory, the gather of the shards and the replicated last rounds must reproduce the
one-GPU proof and the oracle's bytes (multi_composed_sumcheck.rs:64-120), up to the 2^28-entry degree-3 target."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_proofs_match_one_gpu_and_oracle(built, world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ)
    env.setdefault("ZKSC_CHECK_MAX_N", "28" if world == 2 else "26")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + world), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0 and "MULTI-GPU PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def _prove(zk, ctx, t, proto):
    s = t.poly_sum()
    msgs, lens, chal = t.prove(proto, s)
    return s, msgs, lens, chal


@pytest.mark.parametrize("env", [{}, {"ZKSC_NO_TAIL": "1"}, {"ZKSC_GATHER_ENTRIES": "64"}])
def test_single_process_multi_gpu_context(built, env):
    """zksc_ctx_create_multi: ONE context and one host thread over G GPUs of this process (the reference's caller is one process,
    gkr/src/protocol.rs:85).  Sharded tables, the host adds the devices' partial evaluations, one transcript: sums, round messages
    and challenges must equal those of a single-GPU context (which the other tests pin to the oracle), for device-generated tables
    and for host tables uploaded once; also with the resident kernel off (host-side gather) and with an early gather point."""
    import numpy as np
    import zk_cryptography_b200 as zk
    from oracle import cref
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        solo = zk.Context(0)
        for G in (2, 4, 8):
            if _gpu_count() < G:
                continue
            multi = zk.Context(devices=list(range(G)))
            assert multi.devices() == G
            cases = [(3, [2], zk.PROTO_MULTI_PARTIAL), (10, [1], zk.PROTO_SUMCHECK), (12, [2, 3], zk.PROTO_MULTI_PARTIAL), (14, [3], zk.PROTO_COMPOSED),
                     (16, [6], zk.PROTO_MULTI_PARTIAL), (18, [2, 2], zk.PROTO_MULTI_PARTIAL), (22, [3], zk.PROTO_MULTI_PARTIAL), (23, [2], zk.PROTO_MULTI_PARTIAL)]
            for n, degs, proto in cases:
                if n < G.bit_length() - 1:
                    continue
                a, b = zk.Tables.synth(solo, n, degs, 40 + n), zk.Tables.synth(multi, n, degs, 40 + n)
                ra, rb = _prove(zk, solo, a, proto), _prove(zk, multi, b, proto)
                for x, y in zip(ra, rb):
                    assert np.array_equal(x, y), "G=%d n=%d degs=%s: multi-GPU context differs from one GPU" % (G, n, degs)
                # a second proof on the same handles, then the verifier's oracle check (n folds on the devices)
                b.reset()
                rb2 = _prove(zk, multi, b, proto)
                for x, y in zip(ra, rb2):
                    assert np.array_equal(x, y)
                assert np.array_equal(a.evaluate(ra[3]), b.evaluate(ra[3]))
                if n <= 18:      # host tables: uploaded once, every device picks its shard
                    host = a.read_local().reshape(sum(degs), 1 << n, 4)
                    c = zk.Tables.upload(multi, n, degs, [host[i] for i in range(host.shape[0])])
                    rc = _prove(zk, multi, c, proto)
                    for x, y in zip(ra, rc):
                        assert np.array_equal(x, y), "G=%d n=%d: uploaded tables differ" % (G, n)
                    c.free()
                a.free()
                b.free()
            # a batch of independent proofs on the multi-GPU context
            for n, degs, B in ((11, [2], 3), (15, [3], 2)):
                a, b = zk.Tables.synth(solo, n, degs, 60 + n, n_proofs=B), zk.Tables.synth(multi, n, degs, 60 + n, n_proofs=B)
                ra, rb = _prove(zk, solo, a, zk.PROTO_MULTI_PARTIAL), _prove(zk, multi, b, zk.PROTO_MULTI_PARTIAL)
                for x, y in zip(ra, rb):
                    assert np.array_equal(x, y), "G=%d n=%d batch of %d: multi-GPU context differs from one GPU" % (G, n, B)
                a.free()
                b.free()
            # and one case straight against the oracle
            n, degs = 16, [3]
            t = zk.Tables.synth(multi, n, degs, 77)
            s, msgs, lens, chal = _prove(zk, multi, t, zk.PROTO_MULTI_PARTIAL)
            tabs = np.concatenate([cref.synth_table(77, k, n) for k in range(3)])
            osum = cref.poly_sum(n, degs, tabs)
            assert zk.from_mont(s[0]) == osum
            assert (zk._lib.proof_to_bytes(zk.PROTO_MULTI_PARTIAL, msgs[0], lens[0]), zk.from_mont(chal[0])) == cref.prove(2, n, degs, tabs, osum)
            t.free()
            # calls that make no sense on a sharded parent are refused, not mis-executed
            with pytest.raises(zk.ZkscError):
                zk.Tables.alloc(multi, 8, [2])
            multi.close()
        solo.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
