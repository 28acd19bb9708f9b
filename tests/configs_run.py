"""Ad-hoc runs of the larger BASELINE.json configs on the GPU box (not a test):
   python tests/configs_run.py difft | arms15k | chr25k"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tadpole_b200 import Context, TADpole, api
from tadpole_b200.synth import synth_hic, synth_partition_pairs
api.QUIET = True
what = sys.argv[1]
ctx = Context(0)
out = {"config": what}
if what == "difft":
    import torch
    lx, ly = synth_partition_pairs(1000, 15000, 500, seed=1)
    t = time.perf_counter(); res = ctx.difft_batch(lx, ly); out["e2e_first_ms"] = (time.perf_counter() - t) * 1e3
    t = time.perf_counter(); res = ctx.difft_batch(lx, ly); out["e2e_ms"] = (time.perf_counter() - t) * 1e3
    dx, dy = torch.from_numpy(lx).cuda(), torch.from_numpy(ly).cuda()
    do = torch.empty((1000, 15000), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.profile(1)
    for _ in range(5):
        ctx.difft_batch_dev(dx.data_ptr(), dy.data_ptr(), 15000, 1000, do.data_ptr())
    p = ctx.profile(0)["difft"]
    ms = p[0] / p[1]
    out["kernel_ms"] = ms
    out["pairs_per_s"] = 1000 / (ms * 1e-3)
    out["algorithmic_GBps"] = 16 * 15000 * 1000 / (ms * 1e-3) / 1e9
    assert np.array_equal(do.cpu().numpy(), res)
    from oracle import tadpole_oracle as O
    t = time.perf_counter(); ref = O.difft_from_labels_c(lx[0], ly[0]); out["cpu_one_pair_ms"] = (time.perf_counter() - t) * 1e3
    assert (ref == res[0]).all()
elif what in ("arms15k", "chr25k"):
    n = 15000 if what == "arms15k" else 25000
    t = time.perf_counter(); m = synth_hic(n, seed=1, centromere=(what == "arms15k")); out["synth_s"] = time.perf_counter() - t
    for rep in range(2):
        t = time.perf_counter()
        tp = TADpole(m, centromere_search=(what == "arms15k"), ctx=ctx)
        out[f"wall_s_rep{rep}"] = time.perf_counter() - t
        out[f"timings_rep{rep}"] = ctx.timings()
    if what == "arms15k":
        out["p"] = [tp.p.n_pcs, tp.p.optimal_n_clusters]; out["q"] = [tp.q.n_pcs, tp.q.optimal_n_clusters]
        out["merged_tads"] = int(tp.merging_arms.shape[0])
    else:
        out["n_pcs"] = tp.n_pcs; out["optimal_n_clusters"] = tp.optimal_n_clusters; out["levels"] = len(tp.clusters)
print(json.dumps(out), flush=True)
