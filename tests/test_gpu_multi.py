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
