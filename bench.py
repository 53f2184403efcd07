#!/usr/bin/env python
"""bench.py -- sumcheck prover throughput (hypercube evals/s) on N B200s, or the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c1|c4|c4b|c5] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = what the reference's `multi_composed_sumcheck_with_prove_partial_benchmark` times
(sumcheck/benches/multi_composed_sumcheck_benchmark.rs:138-146):
    sum = MultiComposedSumcheckProver::calculate_poly_sum(&poly);  prove_partial(&poly, &sum)
on synthetic seeded tables.  Workloads (BASELINE.json configs):
    c2 (default)  one degree-2 product, 2^24 entries per GPU (2^(24+log2 N) sharded over N GPUs: weak scaling)
    c3            one degree-3 product, 2^28 entries in total, sharded over the N GPUs (strong scaling)
    c1            Sumcheck::prove on one 2^20-entry multilinear (replicated on every GPU)
    c5            64 independent degree-2 proofs of 2^22 entries, 64/N per GPU (replicas, batched launches)
    c4            GKRProtocol::prove on Circuit::random(10): 10 layer sumchecks (largest 2^20 entries) with device-built tables
    c4b           GKRProtocol::prove on a uniform-width layered circuit, width 2^20, depth 8 (linear-time two-phase layer sumchecks)
`value`   : tables resident in HBM when the clock starts (generated on the device).
`e2e`     : the same step through the public API with HOST tables: pinned host -> device copy of every
            table and device -> host copy of the proof inside the timed region.
`--impl reference` : the CPU restatement of the reference (oracle/zkref.c; the Rust reference cannot be
            built in this image) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sumcheck prover hypercube evals/sec"
UNIT = "evals/s"
SEED = 20261017

# name -> (n_vars at N=1, degrees, protocol name, proofs, scaling)
WORKLOADS = {
    "c2": dict(n=24, degs=[2], proto="multi_partial", proofs=1, scaling="weak",
               desc="MultiComposedSumcheckProver: calculate_poly_sum + prove_partial, one degree-2 product, 2^24 entries per GPU"),
    "c3": dict(n=28, degs=[3], proto="multi_partial", proofs=1, scaling="strong",
               desc="MultiComposedSumcheckProver: calculate_poly_sum + prove_partial, one degree-3 product, 2^28 entries in total"),
    "c1": dict(n=20, degs=[1], proto="sumcheck", proofs=1, scaling="replicas",
               desc="Sumcheck: poly_sum + prove, one 2^20-entry multilinear per GPU"),
    "c4": dict(n=20, degs=[2, 2], proto="gkr", proofs=1, scaling="replicas", depth=10,
               desc="GKRProtocol::prove on Circuit::random(10) (the reference's circuit shape: 2^10 inputs, 10 layer sumchecks of 2^2..2^20 entries, "
                    "P=2, d=(2,2); layer tables built on the device), replicated on every GPU"),
    "c4b": dict(n=20, degs=[2, 2], proto="gkr_linear", proofs=1, scaling="replicas", depth=8, log_width=20,
                desc="GKRProtocol::prove on a random layered add/mul circuit of UNIFORM width 2^20, depth 8 (BASELINE config 4 as written; the reference's "
                     "Circuit cannot hold it and its prover would need 2^40-entry layer tables): linear-time two-phase layer sumchecks on 2^20-entry "
                     "tables built from the gate lists (zksc_gkr_prove_linear), replicated on every GPU"),
    "c5": dict(n=22, degs=[2], proto="multi_partial", proofs=64, scaling="strong",
               desc="64 independent prove_partial (degree-2 product, 2^22 entries each), 64/N proofs per GPU, batched launches"),
    "c5s": dict(n=22, degs=[2], proto="multi_partial", proofs=64, scaling="strong", shard_batch=True,
                desc="the same 64 proofs, every one SHARDED over the N GPUs (per-round exchange for all 64 at once, late-round gather of "
                     "the shards inside the resident kernel): BASELINE config 5's gather leg; c5 (replicas) is the faster way to run this batch"),
}


def read_peaks():
    """HBM peak from MEASURED_PEAKS.json (driver-written) else the profiling guide's fallback; the integer
    peak from profiles/int_peak.json (our own IMAD microbenchmark on this pool's B200, tools/ubench.cu)."""
    hbm, src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm, src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    ip = None
    try:
        with open(os.path.join(ROOT, "profiles", "int_peak.json")) as f:
            ip = json.load(f)
    except Exception:
        pass
    return hbm, src, ip


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def algorithmic_bytes(rec):
    """HBM bytes one round-kernel launch must move (DESIGN.md "Algorithmic bytes"): 32 B per element;
    an evaluate-only launch reads both halves of every table (2 * pairs entries); a fused fold+evaluate launch
    reads the previous table (4 * pairs entries) and writes the folded one (2 * pairs)."""
    per_table = (6 if rec["fold"] else 2) * rec["pairs"] * 32
    return per_table * rec["degree"] * rec["proofs"]


def algorithmic_modmuls(rec):
    """Field multiplications one launch needs: (d+1)(d-1) per pair for the d+1 evaluation points (points 0 and 1
    cost d-1 each, as do points 2..d) plus, when the fold is fused in, 2d per pair (both entries of the pair of
    each of the d tables are folded)."""
    d = rec["degree"]
    return ((d + 1) * (d - 1) + (2 * d if rec["fold"] else 0)) * rec["pairs"] * rec["proofs"]


def limb_products(rec):
    """32x32->64 limb products (IMAD.WIDE.U32) one launch executes: fold = 76 per folded entry
    (fixed-multiplicand table; 120 for degrees > 5), a full Montgomery product = 120, the last product of a
    point with d <= 3 stays unreduced = 64.  A fused launch inside prove skips point 1 (derived from the claim)."""
    d = rec["degree"]
    points = d if rec["fold"] else d + 1
    if d == 1:
        per_point = 0
    elif d <= 3:
        per_point = (d - 2) * 120 + 64
    else:
        per_point = (d - 1) * 120
    fold = 2 * d * (76 if d <= 5 else 120) if rec["fold"] else 0
    return (fold + points * per_point) * rec["pairs"] * rec["proofs"]


def proof_digest(proto, msgs, lens, chal):
    """sha256 over the proof bytes (what the reference's transcript absorbs) followed by the challenges, for proof 0."""
    import hashlib
    import numpy as np
    from zk_cryptography_b200 import _lib
    h = hashlib.sha256()
    h.update(_lib.proof_to_bytes(proto, msgs[0], lens[0]))
    h.update(np.ascontiguousarray(chal[0]).tobytes())
    return h.hexdigest()


def round_rows(n, degs, G, round_us, poly_sum_us, hbm_peak, int_tops, gather_at=None):
    """Every round of one proof against its binding roof.  round_us: host wall time per round of zksc_prove (the device pass,
    the transcript step and the bind, so that the rounds add up to the step); round 0's device pass happens inside
    calculate_poly_sum (prove reuses its evaluations), so poly_sum_us is added to round 0.  A sharded proof works on
    2^(n-1-j)/G pairs per rank in round j until the shards are gathered (the last rounds run replicated on all ranks)."""
    rows = []
    lgG = int(math.log2(G))
    for j, us in enumerate(round_us):
        pairs_total = 1 << (n - 1 - j)
        sharded = G > 1 and (n - j) > lgG and (gather_at is None or (1 << (n - j)) > gather_at)
        pairs = pairs_total // G if sharded else pairs_total
        t = us + (poly_sum_us if j == 0 else 0.0)
        t_hbm = t_int = 0.0
        for d in degs:
            rec = {"degree": d, "fold": 1 if j else 0, "pairs": pairs, "proofs": 1}
            t_hbm += algorithmic_bytes(rec) / (hbm_peak * 1e9) * 1e6
            t_int += limb_products(rec) / (int_tops * 1e12) * 1e6
        t_min = max(t_hbm, t_int)
        rows.append({"round": j, "pairs": pairs, "us": round(t, 2), "binding": "hbm" if t_hbm >= t_int else "int32-multiply",
                     "min_us": round(t_min, 3), "frac": round(t_min / t, 4) if t > 0 else None})
    return rows


def parity_check(zk, ctx, rank, G, degs, proto_name, n=20):
    """A proof on the SAME context (sharded when the benchmark's is) against the CPU oracle, outside any timed region:
    hypercube sum, proof bytes and challenges must be identical (oracle/zkref.c = the restatement of
    multi_composed_sumcheck.rs:64-120 / sumcheck.rs:29-61).  Every rank proves; rank 0 compares."""
    from zk_cryptography_b200 import _lib
    proto = {"multi_partial": zk.PROTO_MULTI_PARTIAL, "sumcheck": zk.PROTO_SUMCHECK}[proto_name]
    seed = SEED + 17
    t = zk.Tables.synth(ctx, n, degs, seed)
    s = t.poly_sum()
    msgs, lens, chal = t.prove(proto, s)
    t.free()
    if rank != 0:
        return None
    import numpy as np
    from oracle import cref
    cref.set_threads(cref.max_threads())
    tabs = np.concatenate([cref.synth_table(seed, k, n) for k in range(sum(degs))])
    osum = cref.poly_sum(n, degs, tabs)
    obytes, och = cref.prove({"multi_partial": 2, "sumcheck": 0}[proto_name], n, degs, tabs, osum)
    got = _lib.proof_to_bytes(proto, msgs[0], lens[0])
    ok = zk.from_mont(s[0]) == osum and got == obytes and zk.from_mont(chal[0]) == och
    return {"ok": bool(ok), "n_vars": n, "degrees": degs, "n_gpus": G, "proof_bytes": len(got),
            "against": "oracle/zkref.c prove (sum, proof bytes, challenges byte for byte) on the benchmark's own context"}


def target_c3(zk, torch, dist, ctx, rank, G, args, hbm_peak, int_tops):
    """BASELINE config 3 / the north_star target at every --gpus N: ONE degree-3 product of 2^28 entries in total, sharded over
    the N ranks (strong scaling: the proof is the same for every N, so its hash must be too).  Timed like the headline
    (tables resident, CUDA events on the library's stream, max over ranks), then -- outside the timed region -- every
    rank's proof hash is compared and rank 0 has the ORACLE verify the proof (verifier replay + its own streamed evaluation
    of the 2^28-entry tables at the challenges: oracle/cref.py verify_synth_proof)."""
    import numpy as np
    from zk_cryptography_b200 import _lib
    n, degs, seed = args.target_n, [3], SEED + 3
    tables = zk.Tables.synth(ctx, n, degs, seed)
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device("cuda", torch.cuda.current_device()))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warm = max(3, min(args.steps, 5)), 3
    ps_us, rounds = [], []

    def step(record):
        tables.reset()
        t0 = time.perf_counter()
        s = tables.poly_sum()
        t1 = time.perf_counter()
        out = tables.prove(zk.PROTO_MULTI_PARTIAL, s)
        if record:
            ps_us.append((t1 - t0) * 1e6)
            rounds.append(ctx.round_times())
        return s, out

    for _ in range(warm):
        step(False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    e0.record(stream)
    for _ in range(steps):
        s, (msgs, lens, chal) = step(True)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    if dist is not None:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    sha = proof_digest(zk.PROTO_MULTI_PARTIAL, msgs, lens, chal)
    shas = [sha]
    if dist is not None:
        shas = [None] * G
        dist.all_gather_object(shas, sha)
    tables.free()
    if rank != 0:
        return None
    avg_round = [sum(r[j] for r in rounds) / len(rounds) for j in range(n)]
    rows = round_rows(n, degs, G, avg_round, sum(ps_us) / len(ps_us), hbm_peak, int_tops, gather_at=ctx.gather_entries() if G > 1 else None)
    t_min = sum(r["min_us"] for r in rows)
    out = {"workload": "c3: MultiComposedSumcheckProver calculate_poly_sum + prove_partial, one degree-3 product, 2^%d entries in total, sharded over %d GPU(s)" % (n, G),
           "n_vars": n, "n_gpus": G, "scaling": "strong", "steps": steps, "warmup": warm, "ms": round(ms / steps, 4), "evals_per_s": (1 << n) * steps / (ms * 1e-3),
           "gpu_launches": int(launches), "proof_sha256": sha, "ranks_agree": len(set(shas)) == 1,
           "whole_step": {"min_ms": round(t_min * 1e-3, 4), "frac": round(t_min * 1e-3 / (ms / steps), 4)},
           "rounds_at_or_above_0.60": sum(1 for r in rows if r["frac"] is not None and r["frac"] >= 0.6),
           "per_round": rows, "per_round_clock": "host wall clock inside zksc_prove (device pass + transcript + bind; calculate_poly_sum added to round 0), averaged over the timed steps"}
    if not args.no_cpu:
        from oracle import cref
        cref.set_threads(cref.max_threads())
        t0 = time.perf_counter()
        ok, why = cref.verify_synth_proof(n, 3, seed, zk.from_mont(s[0]), _lib.proof_to_bytes(zk.PROTO_MULTI_PARTIAL, msgs[0], lens[0]), zk.from_mont(chal[0]))
        out["oracle_verified"] = {"ok": bool(ok), "detail": why, "seconds": round(time.perf_counter() - t0, 2), "threads": cref.max_threads(),
                                  "how": "oracle verifier replays the transcript from the proof bytes (challenges, p(0)+p(1) chain), then evaluates the "
                                         "three 2^%d-entry seeded tables at the challenges itself (streamed) and compares with the final claim" % n}
    return out


def run_b200(args):
    import numpy as np
    import torch

    import zk_cryptography_b200 as zk
    from zk_cryptography_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = dict(WORKLOADS[args.workload])
    if wl["proto"] == "gkr":
        return run_gkr(args, wl, world, rank, local_rank, dist)
    if wl["proto"] == "gkr_linear":
        return run_gkr_linear(args, wl, world, rank, local_rank, dist)
    G = world
    lgG = int(math.log2(G))
    ctx = zk.Context(local_rank)
    sharded = wl["scaling"] in ("weak", "strong") and (wl["proofs"] == 1 or wl.get("shard_batch")) and G > 1
    if sharded:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(zk.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(G, rank, bytes(uid.cpu().numpy().tobytes()))
    n = wl["n"] + (lgG if wl["scaling"] == "weak" else 0)
    proofs_total = wl["proofs"]
    proofs_local = (proofs_total if wl.get("shard_batch") else proofs_total // G) if (wl["proofs"] > 1) else 1
    if wl["proofs"] > 1 and proofs_total % G:
        raise SystemExit("proof count must divide by the GPU count")
    proto = {"multi_partial": zk.PROTO_MULTI_PARTIAL, "sumcheck": zk.PROTO_SUMCHECK}[wl["proto"]]
    seed = SEED + (rank * proofs_local if (wl["proofs"] > 1 and not wl.get("shard_batch")) else 0)
    tables = zk.Tables.synth(ctx, n, wl["degs"], seed, n_proofs=proofs_local)
    n_tables, n_evals = tables.n_tables, tables.n_evals

    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device("cuda", local_rank))

    ps_clock = [0.0]

    def step():
        tables.reset()
        t0 = time.perf_counter()
        s = tables.poly_sum()                     # calculate_poly_sum (round-0 pass; its evaluations are reused by prove)
        ps_clock[0] = (time.perf_counter() - t0) * 1e6
        return tables.prove(proto, s)             # prove_partial: n rounds, transcript on the host

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # units per step for the whole job
    if wl["scaling"] == "replicas":
        evals_per_step = (1 << n) * G
    elif wl["proofs"] > 1:
        evals_per_step = (1 << n) * proofs_total
    else:
        evals_per_step = 1 << n

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    ctx.timing(True)
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record(stream)
    acc_rounds, acc_ps = None, 0.0
    for _ in range(args.steps):
        msgs, lens, chal = step()
        rt = ctx.round_times()
        acc_rounds = rt if acc_rounds is None else [a + b for a, b in zip(acc_rounds, rt)]
        acc_ps += ps_clock[0]
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    last_rounds, last_ps_us = [a / args.steps for a in acc_rounds], acc_ps / args.steps
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    recs = ctx.timing_read(cap=1 << 16)
    ctx.timing(False)
    if dist is not None:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    sampler.join(timeout=1.0)
    value = evals_per_step * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (live CUDA events around every launch of the timed region) ----
    hbm_peak, peak_src, int_peak = read_peaks()
    # the integer roof is measured NOW, on this GPU, by the library's own IMAD.WIDE microbenchmark (zksc_int_peak);
    # profiles/int_peak.json (an earlier run of tools/ubench2.cu) is kept next to it for comparison only
    live = ctx.int_peak() / 1e12
    int_peak = {"imad_wide_Tops": round(live, 3), "source": "measured in this run: zksc_int_peak (8 independent mad.wide.u32 per thread, all SMs full)",
                "recorded_earlier": (int_peak or {}).get("imad_wide_Tops")}
    groups = {}
    for r in recs:
        g = groups.setdefault((r["degree"], r["fold"]), dict(ms=0.0, bytes=0, modmuls=0, launches=0, largest=None))
        g["ms"] += r["ms"]; g["bytes"] += algorithmic_bytes(r); g["modmuls"] += algorithmic_modmuls(r); g["launches"] += 1
        if g["largest"] is None or r["pairs"] * r["proofs"] > g["largest"]["pairs"] * g["largest"]["proofs"]:
            g["largest"] = r
    (kd, kf), top = max(groups.items(), key=lambda kv: kv[1]["ms"])
    big = top["largest"]
    big_n = sum(1 for r in recs if (r["degree"], r["fold"]) == (kd, kf) and r["pairs"] == big["pairs"])
    big_ms = sum(r["ms"] for r in recs if (r["degree"], r["fold"]) == (kd, kf) and r["pairs"] == big["pairs"]) / big_n
    achieved = algorithmic_bytes(big) / (big_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
        "traffic": None, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback 6650 GB/s (B200_PROFILING.md)",
        "kernel": "round_kernel<D=%d,FOLD=%s>" % (kd, "true" if kf else "false"),
        "launch": {"pairs_per_table": big["pairs"], "proofs": big["proofs"], "algorithmic_bytes": algorithmic_bytes(big), "avg_ms": round(big_ms, 5),
                   "launches_averaged": big_n},
        "all_launches_of_kernel": {"launches": top["launches"], "ms_per_step": round(top["ms"] / args.steps, 5),
                                   "GBps": round(top["bytes"] / (top["ms"] * 1e-3) / 1e9, 1)},
        "kernel_share_of_step": round(top["ms"] / ms, 4),
    }
    # per-round profile of the timed region: average event-bracketed duration of every launch shape
    prof = {}
    for r in recs:
        k = (r["degree"], r["fold"], r["pairs"], r["proofs"])
        e = prof.setdefault(k, [0.0, 0])
        e[0] += r["ms"]; e[1] += 1
    round_profile = [{"degree": k[0], "fold": k[1], "pairs": k[2], "proofs": k[3], "avg_us": round(v[0] / v[1] * 1e3, 2)}
                     for k, v in sorted(prof.items(), key=lambda kv: -kv[0][2])]
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f).get(roofline["kernel"])
            if tr and tr.get("pairs_per_table") == big["pairs"]:
                roofline["traffic"] = tr["dram_bytes"]
    except Exception:
        pass
    int_roofline = None
    if int_peak:
        prods = limb_products(big) / (big_ms * 1e-3) / 1e12
        int_roofline = {"bound": "int32-multiply", "achieved": round(prods, 3), "peak": int_peak["imad_wide_Tops"], "unit": "T limb-products/s",
                        "frac": round(prods / int_peak["imad_wide_Tops"], 4), "limb_products_per_launch": limb_products(big),
                        "peak_source": int_peak.get("source", "tools/ubench2.cu")}
        # the binding roof of this launch = whichever gives the longer minimum time (north_star: "the slower of the
        # HBM-bandwidth roof and the integer-multiply roof")
        t_hbm = algorithmic_bytes(big) / (hbm_peak * 1e9)
        t_int = limb_products(big) / (int_peak["imad_wide_Tops"] * 1e12)
        roofline["binding"] = {"bound": "hbm" if t_hbm >= t_int else "int32-multiply", "min_ms": round(max(t_hbm, t_int) * 1e3, 5),
                               "frac": round(max(t_hbm, t_int) / (big_ms * 1e-3), 4)}
        # whole proof: every timed launch against its own binding roof
        t_min = sum(max(algorithmic_bytes(r) / (hbm_peak * 1e9), limb_products(r) / (int_peak["imad_wide_Tops"] * 1e12)) for r in recs)
        roofline["whole_step"] = {"min_ms_per_step": round(t_min * 1e3 / args.steps, 5), "frac_of_step": round(t_min * 1e3 / ms, 4),
                                  "note": "launch-timed rounds only; per_round below covers every round"}
    proof_sha = proof_digest(proto, msgs, lens, chal)
    per_round = None
    if proofs_local == 1:
        per_round = round_rows(n, wl["degs"], G if sharded else 1, last_rounds, last_ps_us, hbm_peak, int_peak["imad_wide_Tops"],
                               gather_at=ctx.gather_entries() if sharded else None)
        roofline["whole_step"]["all_rounds_min_ms"] = round(sum(r["min_us"] for r in per_round) * 1e-3, 5)
        roofline["whole_step"]["all_rounds_frac"] = round(sum(r["min_us"] for r in per_round) * 1e-3 / (ms / args.steps), 4)

    # ---- e2e: host tables in pinned memory -> upload -> poly_sum + prove -> proof on the host ----
    e2e = None
    if not args.no_e2e:
        host = torch.empty((proofs_local * n_tables, (1 << n) // (G if sharded else 1), 4), dtype=torch.int64).pin_memory()
        hnp = host.numpy().view(np.uint64)
        hnp[...] = tables.read_local().reshape(hnp.shape)
        views = [hnp[i] for i in range(hnp.shape[0])]
        h2d = hnp.nbytes

        def e2e_step():
            tables.reupload(views, local=sharded)
            s = tables.poly_sum()
            return tables.prove(proto, s)

        for _ in range(max(1, min(args.warmup, 3))):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            m2, l2, c2 = e2e_step()
        f1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms2 = max(f0.elapsed_time(f1), 0.0)
        if dist is not None:
            tms = torch.tensor([ms2], device="cuda", dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms2 = float(tms.item())
        assert np.array_equal(m2, msgs) and np.array_equal(c2, chal), "e2e proof differs from the resident-table proof"
        d2h = int(m2.nbytes + c2.nbytes + l2.nbytes + 32 * proofs_local)
        sequential = {"value": evals_per_step * args.steps / (ms2 * 1e-3), "ms_per_step": round(ms2 / args.steps, 4),
                      "host_wall_ms_per_step": round(wall_ms / args.steps, 4),
                      "api": "zksc_tables_reupload + zksc_poly_sum + zksc_prove, one after the other"}
        # The same K steps double-buffered: the refill of handle B (copy stream) overlaps poly_sum + prove of handle A (compute
        # stream).  Every step still uploads its own inputs and reads its own proof back inside the timed region.
        pair = [tables, zk.Tables.alloc(ctx, n, wl["degs"], n_proofs=proofs_local)]

        def e2e_pipelined(steps):
            pair[0].reupload_begin(views, local=sharded)
            for k in range(steps):
                cur = pair[k % 2]
                cur.reupload_end()
                if k + 1 < steps:
                    pair[(k + 1) % 2].reupload_begin(views, local=sharded)
                res = cur.prove(proto, cur.poly_sum())
            return res

        e2e_pipelined(2)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        t0 = time.perf_counter()
        m3, l3, c3 = e2e_pipelined(args.steps)
        g1.record(stream)
        barrier()
        wall3 = (time.perf_counter() - t0) * 1e3
        ms3 = max(g0.elapsed_time(g1), 0.0)
        if dist is not None:
            tms = torch.tensor([ms3], device="cuda", dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms3 = float(tms.item())
        assert np.array_equal(m3, msgs) and np.array_equal(c3, chal), "pipelined e2e proof differs from the resident-table proof"
        pair[1].free()
        e2e = {"value": evals_per_step * args.steps / (ms3 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
               "ms_per_step": round(ms3 / args.steps, 4), "host_wall_ms_per_step": round(wall3 / args.steps, 4),
               "api": "double-buffered: zksc_tables_reupload_begin(next handle, pinned host tables) on the copy stream || zksc_poly_sum + zksc_prove(current "
                      "handle) on the compute stream, zksc_tables_reupload_end before a handle is proved (C ABI, host buffers); every step uploads its own tables",
               "sequential": sequential}
        if sequential["value"] > e2e["value"]:      # report whichever schedule is faster as the headline, keep both
            e2e = dict(sequential, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=d2h, pipelined={"value": e2e["value"], "ms_per_step": e2e["ms_per_step"]})

    # ---- correctness under the same roof as the numbers (outside every timed region) ----
    parity = None
    if not args.no_cpu and proofs_local == 1:
        parity = parity_check(zk, ctx, rank, G if sharded else 1, wl["degs"], wl["proto"], n=args.parity_n)
    target = None
    if args.workload == "c2" and not args.no_target:
        tables.free()
        target = target_c3(zk, torch, dist, ctx, rank, G, args, hbm_peak, int_peak["imad_wide_Tops"])

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline (oracle port; N=1 only) ----
    cpu = None
    if G == 1 and not args.no_cpu:
        cpu = cpu_baseline(args.workload, sample_n=cpu_sample_n(args))

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"] if wl["scaling"] != "replicas" else "weak",
        "vs_baseline": None, "dtype": "u32 limbs (256-bit modular integer, BLS12-381 Fr, Montgomery)", "data": "synthetic (seeded splitmix64 tables generated on the device)",
        "config": workload_config(args.workload, G, len(_lib.proof_to_bytes(proto, msgs[0], lens[0]))),
        "exchange": ("per round %d field elements per rank, %s" % (n_evals, "through NVLink peer memory inside the round kernels" if ctx.peer_exchange()
                                                                     else "by ncclAllGather + host sum")) if sharded else None,
        "clocks": sampler.summary(), "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "int_roofline": int_roofline,
        "cpu_baseline": cpu, "proof_sha256": proof_sha, "parity": parity, "per_round": per_round, "target_c3": target,
    }
    if args.round_profile:
        out["round_profile"] = round_profile
    print(json.dumps(out))
    sys.stdout.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def gkr_inputs(depth):
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    return [(0x9E3779B97F4A7C15 * (i + 1) + SEED) % R for i in range(1 << depth)]


def run_gkr(args, wl, world, rank, local_rank, dist):
    """c4: one step = GKRProtocol::prove (gkr/src/protocol.rs:21-113) of Circuit::random(depth) on host-resident layer values.
    The unit is the hypercube entries of the layer sumchecks (sum over layers of 2^(2k)).  Inputs live on the host by
    construction (circuit evaluation is the caller's), so `value` and `e2e` time the same call; only the clock differs."""
    import torch

    import zk_cryptography_b200 as zk
    depth = wl["depth"]
    ctx = zk.Context(local_rank)
    zk.set_default_context(ctx)
    circuit = zk.Circuit.random(depth)
    inp = gkr_inputs(depth)
    ev = circuit.evaluation(inp)
    evals_per_step = sum(len(l) ** 2 for l in ev[1:]) * world
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # the caller's data as a Rust caller holds it (&Circuit, &Vec<Vec<Fr>>: Montgomery limbs in host memory); one step = ONE C
    # call, zksc_gkr_prove (layer loop, table construction, layer sumchecks, W evaluations, outer transcript inside the library)
    inst = zk.GKRInstance(circuit, ev)
    for _ in range(args.warmup):
        raw = inst.prove_raw(ctx)
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        raw = inst.prove_raw(ctx)
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    if dist is not None:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    sampler.join(timeout=1.0)
    proof = inst.parse(raw)
    assert zk.GKRProtocol.verify(circuit, inp, proof), "GKR proof does not verify"
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    h2d = sum(2 * 32 * len(l) for l in ev[1:]) + sum(2 * 40 * len(l.layer) for l in circuit.layers)
    d2h = len(proof.to_bytes())
    value = evals_per_step * args.steps / (ms * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import gkrmodel as g
        oc = g.Circuit.random(depth - 2)
        oin = gkr_inputs(depth - 2)
        oev = oc.evaluation(oin)
        t1 = time.perf_counter()
        g.GKRProtocol.prove_sparse(oc, oev, layer_prover=g.c_layer_prover, evaluate=g.c_evaluate)
        dt = time.perf_counter() - t1
        cpu = {"value": sum(len(l) ** 2 for l in oev[1:]) / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "oracle/gkrmodel.py prove_sparse (Python layer driver, layer sumchecks in oracle/zkref.c, 1 thread) on Circuit::random(%d), %.2f s" % (depth - 2, dt)}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (256-bit modular integer, BLS12-381 Fr, Montgomery)",
        "data": "synthetic (Circuit::random gates, seeded inputs)",
        "config": {"workload": "c4: " + wl["desc"], "depth": depth, "layer_sumcheck_entries": [len(l) ** 2 for l in ev[1:]], "degrees": wl["degs"],
                   "l2_policy": "inputs_fit_l2_no_flush (largest layer: 4 x 32 MiB)", "sharding": "replicas" if world > 1 else "single GPU", "proof_bytes": d2h},
        "clocks": sampler.summary(), "gpu_launches": int(launches),
        "e2e": {"value": evals_per_step * args.steps / (wall_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": round(wall_ms / args.steps, 4), "api": "GKRProtocol.prove(circuit, circuit_evaluation): host layer values in, proof out (host wall clock)"},
        "roofline": None, "int_roofline": None, "cpu_baseline": cpu,
        "latency": {"sumcheck_rounds": int(sum(2 * (i + 1) for i in range(depth))), "layers": depth,
                    "us_per_round_incl_layer_setup": round(ms / args.steps * 1e3 / sum(2 * (i + 1) for i in range(depth)), 2)},
        "note": "latency-bound: 110 sumcheck rounds (one host round trip each: the transcript stays on the host) + 10 layer set-ups per proof; "
                "no single dominant kernel, hence a latency object instead of a roofline object",
    }
    print(json.dumps(out))
    sys.stdout.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def layered_inputs(log_width):
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    return [(0xD1B54A32D192ED03 * (i + 1) + SEED) % R for i in range(1 << log_width)]


def layered_units(log_width, depth):
    """hypercube points the linear-time prover evaluates: two phases of 2^k table entries per layer"""
    return depth * 2 * (1 << log_width)


def oracle_layered(lc):
    from oracle import gkrmodel as g
    return g.LayeredCircuit(lc.log_width, [[(int(t), int(a), int(b)) for t, a, b in zip(*layer)] for layer in lc.layers()])


def run_gkr_linear(args, wl, world, rank, local_rank, dist):
    """c4b: one step = GKRProtocol::prove of a uniform-width layered circuit (zksc_gkr_prove_linear) on layer values resident in HBM
    (`value`); `e2e` = Circuit::evaluation from host inputs (zksc_circuit_evaluate: upload + one launch per layer) + the proof,
    host wall clock.  Parity: the same circuit family at width 2^10 against the oracle's DENSE prover (2^20-entry layer tables
    through oracle/zkref.c), byte for byte, and GKRProtocol::verify of the timed proof (wiring polynomials on the device; `--verify-full`
    repeats it in Python integers, about a minute)."""
    import hashlib

    import numpy as np
    import torch

    import zk_cryptography_b200 as zk
    lw, depth = wl["log_width"], wl["depth"]
    ctx = zk.Context(local_rank)
    zk.set_default_context(ctx)
    lc = zk.LayeredCircuit.random([lw] * (depth + 1), SEED, ctx)
    inp = layered_inputs(lw)
    inp_m = zk.to_mont(inp)
    lc.evaluate(inp_m, mont=True)
    units = layered_units(lw, depth) * world
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        raw = lc.prove_raw()
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        raw = lc.prove_raw()
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    rounds_us = ctx.round_times(cap=64)
    # end to end: host inputs -> circuit evaluation on the device -> proof on the host
    e2e_ms = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            lc.evaluate(inp_m, mont=True)
            raw2 = lc.prove_raw()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        assert all(np.array_equal(raw[k], raw2[k]) for k in raw), "end-to-end proof differs from the resident one"
    if dist is not None:
        tms = torch.tensor([ms, e2e_ms or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tms[0].item()), (float(tms[1].item()) if e2e_ms is not None else None)
    sampler.join(timeout=1.0)
    proof = lc.prove()
    proof_bytes = proof.to_bytes()
    sha = hashlib.sha256(proof_bytes).hexdigest()
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    # parity and verification, outside the timed regions
    from oracle import gkrmodel as g
    t0 = time.perf_counter()
    small = zk.LayeredCircuit.random([10] * 4, SEED + 1, ctx)
    sin = layered_inputs(10)
    small.evaluate(sin)
    so = oracle_layered(small)
    want = g.prove_layered(so, so.evaluation(sin), layer_prover=g.c_layer_prover, evaluate=g.c_evaluate)
    parity = {"ok": small.prove().to_bytes() == want.to_bytes(), "what": "width 2^10 x 3 layers of the same circuit family: proof bytes against the oracle's DENSE "
              "prover (oracle/gkrmodel.py prove_layered, 2^20-entry layer tables through oracle/zkref.c)", "seconds": None}
    parity["seconds"] = round(time.perf_counter() - t0, 2)
    assert parity["ok"], "linear-time GKR proof differs from the oracle's dense prover"
    t0 = time.perf_counter()
    verified = {"ok": bool(lc.verify(inp_m, proof)), "what": "GKRProtocol::verify of the timed width-2^%d proof: transcript, round checks and claims on the host, "
                "the wiring polynomials (zksc_circuit_wiring_eval) and the input layer on the device" % lw}
    verified["seconds"] = round(time.perf_counter() - t0, 2)
    if args.verify_full:
        t0 = time.perf_counter()
        verified["python_integers"] = {"ok": bool(lc.verify(inp, proof, device=False)), "what": "the same check with the wiring polynomials in Python integers"}
        verified["python_integers"]["seconds"] = round(time.perf_counter() - t0, 2)
        assert verified["python_integers"]["ok"], "GKR proof does not verify (host-only check)"
    assert verified["ok"], "GKR proof does not verify"
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_gkr_linear_sample(1)
    n_rounds = depth * 2 * lw
    value = units * args.steps / (ms * 1e-3)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (256-bit modular integer, BLS12-381 Fr, Montgomery)",
        "data": "synthetic (seeded random gates and inputs)",
        "config": {"workload": "c4b: " + wl["desc"], "log_width": lw, "depth": depth, "gates": depth << lw, "sumcheck_rounds": n_rounds,
                   "unit_note": "evals = table entries of the two phases of every layer sumcheck (2 x 2^%d x %d per proof); the reference's dense form of "
                                "the same proof sums over 2^%d entries per layer" % (lw, depth, 2 * lw),
                   "l2_policy": "inputs_fit_l2_no_flush (four 32 MiB tables per phase)", "sharding": "replicas" if world > 1 else "single GPU",
                   "proof_bytes": len(proof_bytes)},
        "clocks": sampler.summary(), "gpu_launches": int(launches), "proof_sha256": sha, "parity": parity, "verified": verified,
        "e2e": None if e2e_ms is None else {
            "value": units * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 32 << lw, "d2h_bytes_per_step": int(len(proof_bytes) + (32 << lw)),
            "ms_per_step": round(e2e_ms / args.steps, 4),
            "api": "LayeredCircuit.evaluate(host inputs) + LayeredCircuit.prove_raw(): zksc_circuit_evaluate + zksc_gkr_prove_linear (host wall clock)"},
        "roofline": None, "int_roofline": None, "cpu_baseline": cpu,
        "latency": {"sumcheck_rounds": n_rounds, "layers": depth, "us_per_round_incl_layer_setup": round(ms / args.steps * 1e3 / n_rounds, 2),
                    "last_layer_rounds_us": [round(x, 1) for x in rounds_us[:lw]]},
        "note": "latency-bound: %d sumcheck rounds over tables of at most 2^%d entries (one host round trip each: the transcript stays on the host) "
                "+ 2 table constructions per layer; no single dominant kernel, hence a latency object instead of a roofline object" % (n_rounds, lw),
    }
    print(json.dumps(out))
    sys.stdout.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_gkr_linear_sample(steps, lw=8, depth=4):
    """the reference's (dense) algorithm on the same circuit family at a width it can hold: oracle/gkrmodel.py prove_layered"""
    import random

    from oracle import cref
    from oracle import gkrmodel as g
    th = cref.max_threads()
    cref.set_threads(th)
    rng = random.Random(SEED)
    oc = g.LayeredCircuit([lw] * (depth + 1), [[(rng.randrange(2), rng.randrange(1 << lw), rng.randrange(1 << lw)) for _ in range(1 << lw)] for _ in range(depth)])
    ev = oc.evaluation(layered_inputs(lw))
    t0 = time.perf_counter()
    for _ in range(steps):
        g.prove_layered(oc, ev, layer_prover=g.c_layer_prover, evaluate=g.c_evaluate)
    dt = (time.perf_counter() - t0) / steps
    cref.set_threads(1)
    return {"value": layered_units(lw, depth) / dt, "unit": UNIT, "cores": th, "kind": "port",
            "sample": "oracle/gkrmodel.py prove_layered: the reference's DENSE prover (2^%d-entry layer tables, layer sumchecks in oracle/zkref.c on %d OpenMP "
                      "threads) on a width-2^%d x %d layers circuit of the same family, %.2f s per proof; it cannot hold width 2^20 (2^40-entry tables); "
                      "same unit (two phases of 2^k entries per layer)" % (2 * lw, th, lw, depth, dt)}


def run_reference_gkr_linear(args):
    wl = WORKLOADS[args.workload]
    for _ in range(min(args.warmup, 1)):
        cpu_gkr_linear_sample(1)
    t0 = time.perf_counter()
    cpu = cpu_gkr_linear_sample(args.steps)
    dt = time.perf_counter() - t0
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 limbs (256-bit modular integer, BLS12-381 Fr, Montgomery)", "data": "synthetic (same circuit family and inputs)",
        "config": {"workload": "c4b: " + wl["desc"], "log_width": 8, "depth": 4},
        "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(workload, G, proof_bytes):
    """The `config` object of a line: a function of the workload and the GPU count only, so that both arms print the same one."""
    wl = WORKLOADS[workload]
    lgG = int(math.log2(G))
    sharded = wl["scaling"] in ("weak", "strong") and (wl["proofs"] == 1 or wl.get("shard_batch")) and G > 1
    n = wl["n"] + (lgG if wl["scaling"] == "weak" else 0)
    proofs_local = (wl["proofs"] if wl.get("shard_batch") else wl["proofs"] // G) if wl["proofs"] > 1 else 1
    per_gpu = proofs_local * sum(wl["degs"]) * ((1 << n) // (G if sharded else 1)) * 32
    return {"workload": "%s: %s" % (workload, wl["desc"]), "n_vars": n, "degrees": wl["degs"], "proofs": wl["proofs"], "table_bytes_per_gpu": int(per_gpu),
            "l2_policy": "inputs_exceed_l2" if per_gpu > 126e6 * 2 else "inputs_fit_l2_no_flush",
            "sharding": ("index mod %d (the variables bound last); per-round exchange of the partial evaluations" % G) if sharded else ("replicas" if G > 1 else "single GPU"),
            "proof_bytes": int(proof_bytes)}


def cpu_sample_n(args):
    """n_vars of the CPU sample: the workload's own size for c1 / c2 (a step is ~2 s on 16 threads: same config as the GPU
    arm), a bounded sample (n = 22, stated in the line) for the workloads that would take minutes per step (c3, c5)."""
    if args.cpu_n is not None:
        return args.cpu_n
    return WORKLOADS[args.workload]["n"] if args.workload in ("c1", "c2") else 22


def cpu_workload(workload, sample_n):
    import numpy as np
    from oracle import cref
    wl = WORKLOADS[workload]
    n = min(wl["n"], sample_n)
    D = sum(wl["degs"])
    tabs = np.concatenate([cref.synth_table(SEED, k, n) for k in range(D)])
    proto = {"multi_partial": 2, "sumcheck": 0}[wl["proto"]]
    return cref, wl, n, tabs, proto


def cpu_baseline(workload, sample_n):
    """The oracle port on the host cores: one bounded sample of the same workload at a smaller n
    (cost is linear in 2^n), all OpenMP threads, plus the single-thread figure (the reference itself is
    single-threaded)."""
    cref, wl, n, tabs, proto = cpu_workload(workload, sample_n)
    res = {}
    for label, th in (("all", cref.max_threads()), ("one", 1)):
        cref.set_threads(th)
        nn = n if label == "all" else max(10, n - 3)
        tt = tabs if nn == n else __import__("numpy").concatenate([cref.synth_table(SEED, k, nn) for k in range(sum(wl["degs"]))])
        t0 = time.perf_counter()
        s = cref.poly_sum(nn, wl["degs"], tt)
        cref.prove(proto, nn, wl["degs"], tt, s)
        dt = time.perf_counter() - t0
        res[label] = ((1 << nn) / dt, nn, dt, th)
    return {"value": res["all"][0], "unit": UNIT, "cores": res["all"][3], "kind": "port",
            "sample": "oracle/zkref.c (C restatement with the reference's loop structure), same workload at n_vars=%d (%.2f s), OpenMP over %d threads"
                      % (res["all"][1], res["all"][2], res["all"][3]),
            "single_thread": {"value": res["one"][0], "n_vars": res["one"][1], "seconds": round(res["one"][2], 3)}}


def run_reference(args):
    """The reference's CPU implementation of the path on the host cores.  The reference is Rust and cannot be
    compiled in this image (no cargo/rustc, crates not vendored), so this is the oracle port (`kind: port`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if WORKLOADS[args.workload]["proto"] == "gkr":
        return run_reference_gkr(args)
    if WORKLOADS[args.workload]["proto"] == "gkr_linear":
        return run_reference_gkr_linear(args)
    cref, wl, n, tabs, proto = cpu_workload(args.workload, cpu_sample_n(args))
    th = cref.max_threads()
    cref.set_threads(th)

    def step():
        s = cref.poly_sum(n, wl["degs"], tabs)
        cref.prove(proto, n, wl["degs"], tabs, s)

    pbytes, _ = cref.prove(proto, n, wl["degs"], tabs, cref.poly_sum(n, wl["degs"], tabs)) if args.warmup == 0 else (None, None)
    for _ in range(args.warmup):
        s0 = cref.poly_sum(n, wl["degs"], tabs)
        pbytes, _ = cref.prove(proto, n, wl["degs"], tabs, s0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = (1 << n) * args.steps / dt
    full = n == wl["n"]
    sample = "oracle/zkref.c on %d OpenMP threads; each step = calculate_poly_sum + prove at n_vars=%d (%s)" % (
        th, n, "the workload's own size" if full else "bounded sample of %s; cost is linear in 2^n" % args.workload)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": WORKLOADS[args.workload]["scaling"] if WORKLOADS[args.workload]["scaling"] != "replicas" else "weak",
        "vs_baseline": None, "dtype": "u64 limbs (256-bit modular integer, BLS12-381 Fr, Montgomery)", "data": "synthetic (same seeded tables)",
        "config": workload_config(args.workload, args.gpus, len(pbytes) * (wl["n"] + (int(math.log2(args.gpus)) if wl["scaling"] == "weak" else 0)) // max(n, 1)),
        "cpu_sample_n_vars": n,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": th, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def run_reference_gkr(args):
    """c4 on the host: the oracle's GKR driver (Python) with its layer sumchecks in the C oracle, on a bounded sample
    (Circuit::random(depth - 2): cost is dominated by the largest layers, 4x per extra layer)."""
    from oracle import cref
    from oracle import gkrmodel as g
    wl = WORKLOADS[args.workload]
    depth = wl["depth"] - 2
    th = cref.max_threads()
    cref.set_threads(th)
    oc = g.Circuit.random(depth)
    oev = oc.evaluation(gkr_inputs(depth))
    units = sum(len(l) ** 2 for l in oev[1:])

    def step():
        g.GKRProtocol.prove_sparse(oc, oev, layer_prover=g.c_layer_prover, evaluate=g.c_evaluate)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    sample = "oracle/gkrmodel.py prove_sparse on Circuit::random(%d) (bounded sample of c4), layer sumchecks in oracle/zkref.c on %d OpenMP threads" % (depth, th)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 limbs (256-bit modular integer, BLS12-381 Fr, Montgomery)", "data": "synthetic (same circuit family and inputs)",
        "config": {"workload": "c4: " + wl["desc"], "depth": depth},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": th, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-n", type=int, default=None, help="n_vars of the CPU sample (default: the workload's own n for c1/c2, 22 otherwise)")
    ap.add_argument("--parity-n", type=int, default=20, help="n_vars of the out-of-band parity proof against the oracle")
    ap.add_argument("--target-n", type=int, default=28, help="n_vars of the target_c3 leg (BASELINE config 3: 28)")
    ap.add_argument("--no-target", action="store_true", help="skip the target_c3 leg (degree 3, 2^28 entries, strong scaling)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--verify-full", action="store_true", help="c4b: repeat the verification of the timed width-2^20 proof in Python integers (about a minute)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--round-profile", action="store_true", help="add the per-launch-shape average durations to the JSON line")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
