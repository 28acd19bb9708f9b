"""Ad-hoc (not a test): M = Xc Xc^T by the sliced int8 Gram against the FP64 DMMA Gram -- PCA time and score agreement.
   python tests/mgram_timing.py N [N ...]"""
import sys, os
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
    for name, v in (("fp64-gram", 0), ("sliced-gram", 1024)):
        ctx.set("mgram_min_n", v)
        for rep in range(2):
            ctx.compact(keep); ctx.correlation()
            k = ctx.pca(200)
            t = ctx.timings()
        ctx.profile(1)
        ctx.compact(keep); ctx.correlation(); ctx.pca(200)
        prof = ctx.profile(0)
        out[name] = (ctx.get_scores(keep.size, k), t["pca_ms"], t["pca_applications"], t["pca_iterations"],
                     {kk: (round(vv[0], 3), vv[1]) for kk, vv in prof.items() if vv[1]})
    ref = out["fp64-gram"][0]
    for name, (sc, ms, apps, its, prof) in out.items():
        sgn = np.sign((sc * ref).sum(axis=0)); sgn[sgn == 0] = 1
        err = np.abs(sc * sgn - ref).max() / np.abs(ref).max()
        print(f"N={n} {name:12s} pca_ms={ms:9.2f} applications={apps:.0f} iterations={its:.0f} "
              f"max score diff vs fp64-gram = {err:.2e} profile={prof}", flush=True)
