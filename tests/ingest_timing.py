"""Timing of the input side on the GPU box: python tests/ingest_timing.py N  (not a pytest file).
Device-side parse of an N x N count matrix file against pandas.read_csv + upload of the FP64 matrix."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from tadpole_b200 import Context                     # noqa: E402
from tadpole_b200.synth import synth_hic             # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
m = synth_hic(n, seed=1)
t0 = time.time()
text = "\n".join("\t".join(map(str, row)) for row in m.astype(np.int64).tolist()) + "\n"
path = "/tmp/ingest_%d.tsv" % n
with open(path, "w") as fh:
    fh.write(text)
print(f"N = {n}: text {len(text) / 1e6:.1f} MB ({len(text) / n / n:.2f} B per field), written in {time.time() - t0:.1f} s", flush=True)
ctx = Context(0)
for rep in range(3):
    t0 = time.perf_counter()
    ptr, nn = ctx.ingest_tsv(path)
    wall = (time.perf_counter() - t0) * 1e3
    st = ctx.ingest_stats()
    print(f"  ingest_tsv(file) #{rep}: wall {wall:.1f} ms (library: {st['wall_ms']:.1f} ms, parse kernels {st['parse_ms']:.2f} ms, "
          f"{st['text_bytes'] / st['parse_ms'] / 1e6:.1f} GB/s of text + {8 * n * n / st['parse_ms'] / 1e6:.1f} GB/s written), "
          f"host fields {st['host_fields']}", flush=True)
got = ctx.get_ingested(nn)
assert nn == n and np.array_equal(got, m)
buf = text.encode()
t0 = time.perf_counter()
ctx.ingest_tsv(buf)
print(f"  ingest_tsv(memory): wall {(time.perf_counter() - t0) * 1e3:.1f} ms")
import pandas as pd                                   # noqa: E402
t0 = time.perf_counter()
ref = pd.read_csv(path, sep="\t", header=None, dtype=np.float64).to_numpy()
t1 = time.perf_counter()
bad, _, _ = ctx.filter(ref)
t2 = time.perf_counter()
print(f"  pandas.read_csv: {(t1 - t0) * 1e3:.0f} ms, + upload/filter of the FP64 matrix {(t2 - t1) * 1e3:.0f} ms")
assert np.array_equal(ref, m)
t0 = time.perf_counter()
res = ctx.call(device_ptr=ctx.ingest_tsv(path)[0], n=n, colmajor=0)
print(f"  file -> TADpole result: {(time.perf_counter() - t0) * 1e3:.0f} ms (n_pcs {res['n_pcs']}, {res['n_clusters']} clusters)")
