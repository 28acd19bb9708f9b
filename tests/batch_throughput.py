"""Ad-hoc (not a test): calls/s of tp_call_batch against the number of calls in flight.  python tests/batch_throughput.py [N]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tadpole_b200 import Context
from tadpole_b200.synth import synth_hic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
mats = [synth_hic(n, seed=1 + s) for s in range(16)]
dev = [torch.as_tensor(m, device="cuda") for m in mats]
torch.cuda.synchronize()
ptrs = [d.data_ptr() for d in dev]
ctx = Context(0)
for s in [int(x) for x in os.environ.get('INFLIGHT', '1,2,4,6,8,12,16,24,32').split(',')]:
    ctx.call_batch(None, device_ptrs=ptrs, n=n, inflight=s, tables=False)
    ctx.call_batch(None, device_ptrs=ptrs * 10, n=n, inflight=s, tables=False)
    ms = ctx.last_batch_device_ms
    t0 = time.perf_counter()
    ctx.call_batch(mats * 5, inflight=s)
    wall = time.perf_counter() - t0
    print(f"in flight {s:2d}: device-resident {160 / ms * 1e3:7.1f} calls/s   host matrices + tables {80 / wall:7.1f} calls/s", flush=True)
