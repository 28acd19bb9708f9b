"""Ad-hoc (not a test): the subspace-iteration log of one call (TADPOLE_DEBUG=1 python tests/pca_log.py N [N ...])."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tadpole_b200 import Context, api
from tadpole_b200.synth import synth_hic, synth_hic_gpu
api.QUIET = True
ctx = Context(0)
for n in [int(a) for a in sys.argv[1:]] or [2000]:
    m = synth_hic(n, seed=1) if n <= 8000 else synth_hic_gpu(n, seed=3).cpu().numpy()
    ctx.call(m)
    print(f"---- n = {n}", file=sys.stderr, flush=True)
    ctx.profile(1)
    r = ctx.call(m)
    p = ctx.profile(0)
    print({k: (round(v[0], 3), v[1]) for k, v in p.items() if v[1]}, ctx.timings(), file=sys.stderr, flush=True)
