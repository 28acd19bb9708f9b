"""Ad-hoc stage timing on the GPU box (not a test): python tests/stage_timing.py [N]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tadpole_b200 import Context
from tadpole_b200.synth import synth_hic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = Context(0)
m = synth_hic(n, seed=1)
for rep in range(4):
    t = time.perf_counter()
    r = ctx.call(m)
    wall = (time.perf_counter() - t) * 1e3
    tm = ctx.timings()
    print(f"N={n} rep={rep} wall={wall:.2f} ms n_pcs={r['n_pcs']} ncl={r['n_clusters']} nf={r['nf']} "
          + " ".join(f"{k}={v:.3f}" for k, v in tm.items()), flush=True)
ctx.profile(1)
r = ctx.call(m)
prof = ctx.profile(0)
print("profile (ms, launches):", {k: (round(v[0], 3), v[1]) for k, v in prof.items()})
print("launches", ctx.launches)
