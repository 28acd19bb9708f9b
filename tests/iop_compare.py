"""Ad-hoc (not a test): accuracy and time of the sliced int8 PCA operator against the pure FP64 DMMA solve.
   python tests/iop_compare.py N [N ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tadpole_b200 import Context
from tadpole_b200.synth import synth_hic
ctx = Context(0)
for n in [int(a) for a in sys.argv[1:]]:
    m = synth_hic(n, seed=1)
    bad, _, _ = ctx.filter(m)
    keep = np.flatnonzero(~bad).astype(np.int32)
    out = {}
    for name, cfg in (("fp64", dict(iop_min_n=0)), ("x5+fp64", dict(iop_min_n=1024, iop_final=0)),
                      ("x5+x8", dict(iop_min_n=1024, iop_final=8))):
        for k_, v in cfg.items():
            ctx.set(k_, v)
        for rep in range(2):
            ctx.compact(keep); ctx.correlation()
            k = ctx.pca(200)
            t = ctx.timings()
        out[name] = (ctx.get_scores(keep.size, k), t["pca_ms"], t["pca_applications"], t["pca_iterations"])
    ref = out["fp64"][0]
    for name, (sc, ms, apps, its) in out.items():
        sgn = np.sign((sc * ref).sum(axis=0)); sgn[sgn == 0] = 1
        err = np.abs(sc * sgn - ref).max() / np.abs(ref).max()
        print(f"N={n} {name:8s} pca_ms={ms:9.2f} applications={apps:.0f} iterations={its:.0f} max score diff vs fp64 = {err:.2e}", flush=True)
