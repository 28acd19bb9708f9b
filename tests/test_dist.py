"""world_size-2 test of the multi-GPU host logic on CPU (gloo): candidates are sharded
rank-interleaved, each rank fills its rows of the score matrix, rows are combined with an
all-reduce, and every rank selects the same optimum as a single-rank run."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import tadpole_oracle as O
    from tadpole_b200 import sharding
    from tadpole_b200.synth import synth_hic
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = synth_hic(120, seed=7)
    lm = O.load_mat_numeric(m)
    pcs = O.prcomp_scores(O.sparse_cor(lm.mat), 24)
    k = pcs.shape[1]

    def sweep(cand_begin, cand_stride):            # stands in for Context.sweep on this rank's GPU
        rows = {}
        for c in range(cand_begin, k, cand_stride):
            rows[c] = O.candidate_scores(pcs, c + 1, 2)[0]
        width = max(len(v) for v in rows.values())
        sc = np.full((k, width), np.nan)
        ncl = np.zeros(k, dtype=np.int32)
        for c, v in rows.items():
            sc[c, : len(v)] = v
            ncl[c] = len(v)
        return ncl, sc

    ncl, scores = sharding.sharded_sweep(sweep, k, rank, world)
    opt = sharding.select(scores)
    if rank == 0:
        np.save(out, np.concatenate(([opt[0], opt[1]], ncl)))
    dist.destroy_process_group()


def test_sharded_sweep_world2(tmp_path):
    import torch.multiprocessing as mp
    from oracle import tadpole_oracle as O
    from tadpole_b200.synth import synth_hic
    out = str(tmp_path / "r.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    m = synth_hic(120, seed=7)
    lm = O.load_mat_numeric(m)
    pcs = O.prcomp_scores(O.sparse_cor(lm.mat), 24)
    per = [O.candidate_scores(pcs, i, 2) for i in range(1, 25)]
    scores, opcs, ok = O.reduce_scores([p[0] for p in per])
    assert (int(got[0]), int(got[1])) == (opcs, ok)
    assert got[2:].astype(int).tolist() == [p[1] for p in per]


def test_candidate_partition_is_balanced():
    from tadpole_b200 import sharding
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += list(range(*sharding.candidate_range(r, world, 200)))
        assert sorted(seen) == list(range(200))


def _env_worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import json
    import torch.distributed as dist
    from tadpole_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env = sharding.DistEnv(ctx=None, init_nccl=False)          # host-side plan only: no GPU here
    got = env.exchange((env.my_arm, {"rank": rank, "n_pcs": 10 + rank}))
    with open(os.path.join(outdir, f"r{rank}.json"), "w") as fh:
        json.dump({"arm": env.my_arm, "arm_ranks": env.arm_ranks, "got": got}, fh)
    dist.destroy_process_group()


def test_arm_plan_and_exchange_world2(tmp_path):
    """Arms are dealt out to halves of the job and the per-arm result summaries reach every rank."""
    import json
    import torch.multiprocessing as mp
    from tadpole_b200 import sharding
    assert sharding.arm_plan(1) == {"p": [0], "q": [0]}
    assert sharding.arm_plan(2) == {"p": [0], "q": [1]}
    assert sharding.arm_plan(8) == {"p": [0, 1, 2, 3], "q": [4, 5, 6, 7]}
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_env_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = json.load(open(tmp_path / "r0.json"))
    r1 = json.load(open(tmp_path / "r1.json"))
    assert (r0["arm"], r1["arm"]) == ("p", "q") and r0["arm_ranks"] == [0] and r1["arm_ranks"] == [1]
    assert r0["got"] == r1["got"] == [["p", {"rank": 0, "n_pcs": 10}], ["q", {"rank": 1, "n_pcs": 11}]]
