"""Ad-hoc (not a test): calls/s at N bins with 1, 2, 4, 8 calls in flight on one GPU.  python tests/batch_throughput.py [N]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tadpole_b200 import ContextPool
from tadpole_b200.synth import synth_hic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
mats = [synth_hic(n, seed=1 + i) for i in range(4)]
for streams in (1, 2, 4, 8):
    pool = ContextPool(0, streams)
    work = [mats[i % 4] for i in range(8 * streams)]
    fn = lambda ctx, m: ctx.call(m, want_scores=True)["n_pcs"]
    pool.map(fn, work[: 2 * streams])                 # warm-up: buffers allocated
    for c in pool.contexts: c.sync()
    t = time.perf_counter(); pool.map(fn, work)
    for c in pool.contexts: c.sync()
    dt = time.perf_counter() - t
    print(f"N={n} streams={streams} calls={len(work)} {len(work) / dt:.1f} calls/s ({1e3 * dt / len(work):.2f} ms per call amortised)", flush=True)
    pool.close()
