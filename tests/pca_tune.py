"""Ad-hoc tuning of the PCA driver on the GPU box (not a test)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tadpole_b200 import Context
from tadpole_b200.synth import synth_hic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = Context(0)
m = synth_hic(n, seed=1)
for inner in (2, 3, 4):
    for blk in (0, 224, 288):
        ctx.set("pca_inner", inner); ctx.set("pca_block", blk)
        try:
            for rep in range(2):
                r = ctx.call(m)
            tm = ctx.timings()
            print(f"inner={inner} block={blk} total={tm['total_ms']:.2f} pca={tm['pca_ms']:.2f} its={tm['pca_iterations']:.0f} "
                  f"apps={tm['pca_applications']:.0f} sweeps={tm['jacobi_sweeps']:.0f} n_pcs={r['n_pcs']} ncl={r['n_clusters']}", flush=True)
        except Exception as e:
            print(f"inner={inner} block={blk} FAILED {e}", flush=True)
