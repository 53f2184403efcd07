"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints ONE JSON line with the keys the driver
reads, on rank 0 only, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_line(built):
    r = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-n", "14"])
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sumcheck prover hypercube evals/sec" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c2")


@pytest.mark.parametrize("workload", ["c4", "c4b"])
def test_reference_arm_gkr_workloads(built, workload):
    """the GKR workloads' reference arm: the oracle's (dense) GKR driver on a bounded sample of the same circuit family"""
    r = run_bench(["--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-1500:]
    d = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["value"] > 0
    assert d["config"]["workload"].startswith(workload + ":") and d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_stay_silent(built):
    r = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-n", "12"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_ignores_torchrun_thread_cap(built):
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use every CPU of the affinity mask"""
    r = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-n", "12"], {"OMP_NUM_THREADS": "1"})
    d = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_product_arm_needs_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run_bench(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
