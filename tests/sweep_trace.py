"""Ad-hoc (not a test): cycle breakdown of one CONISS merge step.  TADPOLE_SWEEP_TRACE=1 python tests/sweep_trace.py N [N ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tadpole_b200 import Context
ctx = Context(0)
rng = np.random.default_rng(0)
for n in [int(a) for a in sys.argv[1:]] or [2000]:
    # smooth + noise score columns with a decaying scale, like PC scores
    s = np.cumsum(rng.standard_normal((n, 200)), axis=0) * (1.0 + np.arange(200)) ** -0.5
    ctx.set_scores(s)
    for rep in range(2):
        ctx.profile(1)
        ctx.sweep(200)
        p = ctx.profile(0)
    print(f"n={n}: sweep kernel {p['coniss_sweep'][0]:.3f} ms, {200 * (n - 1) / p['coniss_sweep'][0] / 1e3:.1f} M merges/s", file=sys.stderr, flush=True)
