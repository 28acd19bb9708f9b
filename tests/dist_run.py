"""Ad-hoc multi-GPU run (not a test), one process per GPU:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
       tests/dist_run.py [check|time] N_BINS [centromere] [fast]
check: rank 0 also runs the same call on its GPU alone and the results must be identical.
time: two timed collective calls, stage timings of rank 0.
fast: the matrix is drawn on the GPU (torch.poisson on the same lambda structure, same seed on every rank, checksum
compared across ranks) instead of numpy's generator, which needs 26 s per process at 25 000 bins."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from tadpole_b200 import Context, TADpole, api, sharding
from tadpole_b200.synth import synth_hic

mode, n = sys.argv[1], int(sys.argv[2])
cen = "centromere" in sys.argv[3:]
fast = "fast" in sys.argv[3:]
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
api.QUIET = True
ctx = Context(lr)
env = sharding.DistEnv(ctx)
def synth_fast(n, seed=1):
    from tadpole_b200.synth import _random_blocks
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda", lr)
    idx = torch.arange(n, device=dev)
    blocks = [(torch.from_numpy(_random_blocks(rng, n, n / 5.0 if ml is None else float(ml))).to(dev), bo)
              for ml, bo in ((None, 1.5), (40, 3.0), (10, 6.0))]
    g = torch.Generator(device=dev); g.manual_seed(seed)
    mat = torch.empty((n, n), dtype=torch.float64, device=dev)
    step = max(1, (1 << 26) // n)
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        lam = 200.0 / ((idx[r0:r1, None] - idx[None, :]).abs().float() + 1.0)
        for ids, bo in blocks:
            lam = torch.where(ids[r0:r1, None] == ids[None, :], lam * bo, lam)
        mat[r0:r1] = torch.poisson(lam, generator=g).double()
    mat = torch.triu(mat) + torch.triu(mat, 1).T
    z = torch.from_numpy(rng.choice(n, size=int(round(0.005 * n)), replace=False)).to(dev)
    mat[z, :] = 0.0; mat[:, z] = 0.0
    host = mat.cpu().numpy()
    del mat, lam; torch.cuda.empty_cache()
    return host

m = synth_fast(n) if fast else synth_hic(n, seed=1, centromere=cen)
out = {"mode": mode, "n": n, "world": world, "centromere": cen, "fast_synth": fast}
if fast:
    sums = env.exchange(float(m.sum()))
    assert all(v == sums[0] for v in sums), "ranks drew different matrices"
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    tp = TADpole(m, centromere_search=cen, ctx=ctx, dist=env)
    ctx.sync(); dist.barrier()
    out[f"wall_s_rep{rep}"] = time.perf_counter() - t
    out[f"timings_rep{rep}"] = ctx.timings()
ctx.profile(1)
tp = TADpole(m, centromere_search=cen, ctx=ctx, dist=env)
out["profile_ms"] = {k: round(v[0], 3) for k, v in ctx.profile(0).items() if v[1]}
if cen:
    out["p"] = [tp.p.n_pcs, tp.p.optimal_n_clusters]; out["q"] = [tp.q.n_pcs, tp.q.optimal_n_clusters]
    out["merged_tads"] = int(tp.merging_arms.shape[0])
else:
    out["n_pcs"] = tp.n_pcs; out["optimal_n_clusters"] = tp.optimal_n_clusters
# every rank must hold the same object
summ = (tp.merging_arms.tobytes() if cen else (tp.n_pcs, tp.optimal_n_clusters, tp.scores.tobytes(), tp.dendro.seqdist.tobytes()))
every = env.exchange(summ)
out["all_ranks_identical"] = all(e == every[0] for e in every)
if mode == "check":
    if rank == 0:
        ctx.comm_select(-1)
        ref = TADpole(m, centromere_search=cen, ctx=ctx)
        if cen:
            out["same_as_single_gpu"] = bool(np.array_equal(ref.merging_arms, tp.merging_arms)
                                             and ref.p.n_pcs == tp.p.n_pcs and ref.q.n_pcs == tp.q.n_pcs)
        else:
            out["same_as_single_gpu"] = bool(ref.n_pcs == tp.n_pcs and ref.optimal_n_clusters == tp.optimal_n_clusters
                                             and np.array_equal(ref.dendro.seqdist, tp.dendro.seqdist)
                                             and np.array_equal(ref.scores, tp.scores, equal_nan=True))
            out["max_score_diff"] = float(np.nanmax(np.abs(ref.scores - tp.scores)))
        out["single_gpu_timings"] = ctx.timings()
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
ctx.close()
dist.destroy_process_group()
