"""Ad-hoc multi-GPU run (not a test), one process per GPU:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
       tests/dist_run.py [check|time] N_BINS [centromere]
check: rank 0 also runs the same call on its GPU alone and the results must be identical.
time: two timed collective calls, stage timings of rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from tadpole_b200 import Context, TADpole, api, sharding
from tadpole_b200.synth import synth_hic

mode, n = sys.argv[1], int(sys.argv[2])
cen = len(sys.argv) > 3 and sys.argv[3] == "centromere"
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
api.QUIET = True
ctx = Context(lr)
env = sharding.DistEnv(ctx)
m = synth_hic(n, seed=1, centromere=cen)
out = {"mode": mode, "n": n, "world": world, "centromere": cen}
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
if mode == "check":
    # every rank must hold the same object
    summ = (tp.merging_arms.tobytes() if cen else (tp.n_pcs, tp.optimal_n_clusters, tp.scores.tobytes(), tp.dendro.seqdist.tobytes()))
    every = env.exchange(summ)
    out["all_ranks_identical"] = all(e == every[0] for e in every)
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
