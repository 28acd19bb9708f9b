"""Wall time of one TADpole() call from a dense host matrix against the same call from its upper-triangle pixels
(SparseCounts -> tp_ingest_coo).  Usage: python tests/coo_timing.py [bins ...]   (a tool, not a test)"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tadpole_b200 import Context, SparseCounts, TADpole  # noqa: E402
from tadpole_b200.synth import synth_hic_gpu  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [2000, 8000, 25000]
    ctx = Context(0)
    out = []
    for n in sizes:
        d = synth_hic_gpu(n, seed=7, device=0)
        up = torch.triu(d)
        idx = up.nonzero()
        b1 = idx[:, 0].to(torch.int32).cpu().numpy()
        b2 = idx[:, 1].to(torch.int32).cpu().numpy()
        v = up[idx[:, 0], idx[:, 1]].cpu().numpy()
        m = d.cpu().numpy()
        del d, up, idx
        torch.cuda.empty_cache()
        src = SparseCounts(b1, b2, v, n)
        res = {}
        for name, arg in (("dense", m), ("pixels", src)):
            ts = []
            for rep in range(3):
                t0 = time.perf_counter()
                r = TADpole(arg, ctx=ctx)
                ts.append((time.perf_counter() - t0) * 1e3)
            res[name] = dict(wall_ms=min(ts), n_pcs=r.n_pcs, n_clusters=r.optimal_n_clusters, seq=r.dendro.seqdist)
            if name == "pixels":
                res[name]["ingest"] = ctx.ingest_stats()
        same = bool(np.array_equal(res["dense"]["seq"], res["pixels"]["seq"]) and res["dense"]["n_pcs"] == res["pixels"]["n_pcs"])
        line = dict(bins=n, pixels=int(b1.size), density_upper=round(b1.size / (n * (n + 1) / 2), 4),
                    dense_bytes_uploaded=int(n * (n + 1) // 2 * 8), pixel_bytes_uploaded=int(b1.size * 16),
                    dense_wall_ms=round(res["dense"]["wall_ms"], 2), pixels_wall_ms=round(res["pixels"]["wall_ms"], 2),
                    ingest_wall_ms=round(res["pixels"]["ingest"]["wall_ms"], 2), ingest_device_ms=round(res["pixels"]["ingest"]["parse_ms"], 3),
                    same_result=same)
        print(json.dumps(line), flush=True)
        out.append(line)
    ctx.close()


if __name__ == "__main__":
    main()
