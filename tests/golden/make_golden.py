"""Generates the committed golden fixtures from the CPU oracle (run here, in the build container).

  difft_control_case.json  -- the reference's own fixture pair /root/reference/inst/extdata/
                              {control,case}.bed through the oracle's literal restatement of
                              R/DiffT.R; the values are pinned against the numbers read off the
                              reference's misc/DiffT_score.png (L = 194,
                              un-normalised total 1777, 0.064716@23, 0.881823@174, ...).
  pipeline_n160.json       -- a seeded 160-bin synthetic matrix and the oracle's TADpole() result.
  pipeline_n1100.json      -- the oracle's TADpole(max_pcs = 40) result for golden_int_matrix(1100, 8) (intgen.py: integer-only
                              arithmetic, so the matrix is reproduced bit for bit on any machine): a size that takes the
                              subspace iteration and the tcgen05 int8 kernels on the GPU (Nf = 1089 >= 1024).  The matrix is
                              not stored: (n, seed) and the SHA-256 of its bytes are.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import tadpole_oracle as O  # noqa: E402
from tadpole_b200.synth import synth_hic  # noqa: E402
sys.path.insert(0, HERE)
from intgen import golden_int_matrix  # noqa: E402

REF = "/root/reference/inst/extdata"


def main():
    control = O.read_bed(os.path.join(REF, "control.bed"))
    case = O.read_bed(os.path.join(REF, "case.bed"))
    raw = O.difft(control, case, raw=True)
    norm = O.difft(control, case)
    assert raw.size == 194 and raw[-1] == 1777
    with open(os.path.join(HERE, "difft_control_case.json"), "w") as fh:
        json.dump(dict(control=control.tolist(), case=case.tolist(), raw_cumulative=raw.tolist(),
                       normalised=[float(v) for v in norm]), fh)
    m = synth_hic(160, seed=3)
    r = O.tadpole(m)
    with open(os.path.join(HERE, "pipeline_n160.json"), "w") as fh:
        json.dump(dict(matrix=m.astype(int).tolist(), n_pcs=r.n_pcs, optimal_n_clusters=r.optimal_n_clusters,
                       clusters={str(k): v.tolist() for k, v in r.clusters.items()},
                       scores=[[None if np.isnan(x) else float(x) for x in row] for row in r.scores],
                       seqdist=[float(v) for v in r.seqdist]), fh)
    n, seed, max_pcs = 1100, 8, 40
    m = golden_int_matrix(n, seed)
    r = O.tadpole(m, max_pcs=max_pcs)
    with open(os.path.join(HERE, "pipeline_n1100.json"), "w") as fh:
        json.dump(dict(n=n, seed=seed, max_pcs=max_pcs, matrix_sha256=hashlib.sha256(np.ascontiguousarray(m).tobytes()).hexdigest(),
                       n_pcs=r.n_pcs, optimal_n_clusters=r.optimal_n_clusters,
                       clusters={str(k): v.tolist() for k, v in r.clusters.items()},
                       scores=[[None if np.isnan(x) else float(x) for x in row] for row in r.scores],
                       seqdist=[float(v) for v in r.seqdist]), fh)


if __name__ == "__main__":
    main()
