"""Ad-hoc ncu target (not a test): python tests/prof_one.py [N] [calls] -- `calls` TADpole pipeline calls on one
synthetic N-bin matrix, nothing else on the GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tadpole_b200 import Context
from tadpole_b200.synth import synth_hic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = Context(0)
m = synth_hic(n, seed=1)
for rep in range(calls):
    r = ctx.call(m)
    print(n, rep, r["n_pcs"], r["n_clusters"], ctx.launches, flush=True)
