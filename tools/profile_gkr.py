"""cProfile of one GKRProtocol.prove on Circuit.random(10) (where does the host time go?)"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_cryptography_b200 as zk
from bench import gkr_inputs
ctx = zk.Context(0); zk.set_default_context(ctx)
c = zk.Circuit.random(10); ev = c.evaluation(gkr_inputs(10))
for _ in range(3): zk.GKRProtocol.prove(c, ev, ctx)
t0 = time.perf_counter(); zk.GKRProtocol.prove(c, ev, ctx); print("prove: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
inst = zk.GKRInstance(c, ev)
for _ in range(3): inst.prove_raw(ctx)
t0 = time.perf_counter(); inst.prove_raw(ctx); print("prove_raw (C call only): %.2f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter(); zk.GKRProtocol.prove_layerwise(c, ev, ctx); print("prove_layerwise: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); zk.GKRProtocol.prove(c, ev, ctx); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
