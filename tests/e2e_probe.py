"""Ad-hoc (not a test): where the end-to-end time of a batch goes (python tests/e2e_probe.py)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tadpole_b200 import Context, TADpole_batch, api
from tadpole_b200.synth import synth_hic
api.QUIET = True
n, B = 2000, 16
host = []
for s in range(B):
    t = torch.empty((n, n), dtype=torch.float64, pin_memory=True); t.numpy()[:] = synth_hic(n, seed=1 + s); host.append(t)
hn = [t.numpy() for t in host]
dev = [t.cuda() for t in host]; torch.cuda.synchronize()
ctx = Context(0)
def timeit(f, reps=5):
    f(); t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e3
print("device ptrs, no tables     ms/step", timeit(lambda: ctx.call_batch(None, device_ptrs=[d.data_ptr() for d in dev], n=n, inflight=8, tables=False)), "device_ms", ctx.last_batch_device_ms)
print("pinned host, no tables     ms/step", timeit(lambda: ctx.call_batch(hn, inflight=8, tables=False)), "device_ms", ctx.last_batch_device_ms)
print("pinned host, tables        ms/step", timeit(lambda: ctx.call_batch(hn, inflight=8, tables=True)), "device_ms", ctx.last_batch_device_ms)
def percall(**kw):
    r = ctx.call_batch(**kw)
    return float(np.mean([x["device_ms"] for x in r])), float(np.max([x["device_ms"] for x in r]))
print("per-call device ms (mean, max): device ptrs", percall(mats=None, device_ptrs=[d.data_ptr() for d in dev], n=n, inflight=8, tables=False),
      " pinned host", percall(mats=hn, inflight=8, tables=False))
for infl in (1, 2, 4):
    a = timeit(lambda: ctx.call_batch(None, device_ptrs=[d.data_ptr() for d in dev], n=n, inflight=infl, tables=False))
    b = timeit(lambda: ctx.call_batch(hn, inflight=infl, tables=False))
    print(f"inflight {infl}: device ptrs {a:.1f} ms/step, pinned host {b:.1f} ms/step")
print("TADpole_batch              ms/step", timeit(lambda: TADpole_batch(hn, ctx=ctx, streams=8)))
for infl in (4, 6, 8, 12):
    print(f"TADpole_batch inflight {infl:2d}   ms/step", timeit(lambda: TADpole_batch(hn, ctx=ctx, streams=infl), reps=8))
ctx.set("sync_blocking", 1)
for infl in (8, 12):
    print(f"blocking sync, inflight {infl:2d} ms/step", timeit(lambda: TADpole_batch(hn, ctx=ctx, streams=infl), reps=8))
ctx.set("sync_blocking", 0)
B2 = hn + hn
print("TADpole_batch 32 matrices  ms/16 calls", timeit(lambda: TADpole_batch(B2, ctx=ctx, streams=8)) / 2)
pr = cProfile.Profile(); pr.enable(); TADpole_batch(hn, ctx=ctx, streams=8); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
