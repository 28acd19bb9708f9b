#!/usr/bin/env python
"""bench.py -- TADpole calls/sec on synthetic Hi-C matrices (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

A "step" is one full TADpole() call on one synthetic matrix: bad-column filter, Pearson
correlation, PCA to max_pcs = 200 components, the n_pcs sweep (CONISS per candidate), broken
stick + Calinski-Harabasz, selection.  Workload at N = 1: BASELINE.json configs[1], a 2,000-bin
matrix with nested block TADs and power-law decay.  With N > 1 every rank calls its own
matrices (multi-chromosome batch, no data-path collective): weak scaling.

value  : calls/s with the input matrices already resident in HBM (device pointers through the
         C ABI), timed with CUDA events on the library's streams, max over ranks.  A single 2000-bin call
         leaves most of the GPU idle (its b x b eigen / Cholesky kernels run on one 8-CTA cluster), so, as a
         genome-wide run over many chromosomes would, --streams S independent calls are in flight per GPU
         (tadpole_b200.batch: one context, stream and host thread each); the K steps are dealt out to them.
         `single_call` in the JSON line is the same measured with one call at a time (latency), and the
         per-kernel rooflines are taken from that pass.
e2e    : the same through the public API TADpole(matrix) with the matrix in pinned HOST memory;
         H2D copy of the matrix and D2H of the results are inside the timed region.
A pool of different matrices larger than L2 (5 x 32 MB) is cycled, so no step re-reads a warm input.

The reference is a pure-R package and R is not installed here or on the GPU box, so
`--impl reference` and `cpu_baseline` time oracle/ (numpy + C restatement in the reference's
algorithmic shape: full LAPACK SVD, per-candidate O(N^2) dist + Lance-Williams CONISS, per-level
Calinski-Harabasz) on the host cores, on a bounded sample of candidates extrapolated to all 200.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_BINS = 2000
MAX_PCS = 200
POOL = 5
METRIC = "tadpole_calls_per_sec"
UNIT = "calls/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--streams", type=int, default=8,
                    help="independent calls in flight per GPU (own context / stream / host thread each)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bins", type=int, default=N_BINS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sync-blocking", default="auto", choices=["auto", "0", "1"],
                    help="host threads sleep on a blocking-sync event while they wait for the GPU instead of spinning in "
                         "cudaStreamSynchronize; auto = when ranks x calls in flight exceed half of the host's logical CPUs")
    ap.add_argument("--large-n", type=int, default=8000,
                    help="bins of one extra profiled call whose tensor-kernel rooflines are reported beside the workload's "
                         "(N = 1 only; 0 = skip)")
    ap.add_argument("--cpu-sample", type=int, default=8, help="candidates timed per CPU sample")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle, timed (the only place besides tests/ and smoke() that touches oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_sample(mat, ncand_sample, threads):
    """One bounded sample of the reference-shaped CPU path.  Returns (estimated seconds for the
    full call, detail dict)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import tadpole_oracle as O
    t0 = time.perf_counter()
    lm = O.load_mat_numeric(mat)
    cor = O.sparse_cor(lm.mat)
    k = min(MAX_PCS, lm.mat.shape[0])
    pcs = O.prcomp_scores(cor, k)
    t_front = time.perf_counter() - t0
    cands = np.unique(np.linspace(1, k, ncand_sample).round().astype(int))
    t1 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:      # foreach %dopar% over candidates
        list(ex.map(lambda i: O.candidate_scores(pcs, int(i), 2), cands))
    t_sweep = time.perf_counter() - t1
    est = t_front + t_sweep * (k / len(cands))
    return est, dict(front_s=round(t_front, 3), sweep_sample_s=round(t_sweep, 3), candidates=len(cands), k=int(k))


def run_reference(args, rank, world):
    if rank != 0:
        return
    from tadpole_b200.synth import synth_hic
    threads = os.cpu_count() or 1
    mats = [synth_hic(args.bins, seed=1 + s) for s in range(min(POOL, max(1, args.steps)))]
    for w in range(min(args.warmup, 1)):
        cpu_sample(mats[0], 2, threads)
    ests, detail = [], None
    t0 = time.perf_counter()
    for s in range(args.steps):
        est, detail = cpu_sample(mats[s % len(mats)], args.cpu_sample, threads)
        ests.append(est)
    wall = time.perf_counter() - t0
    sec = float(np.mean(ests))
    val = 1.0 / sec
    sample = (f"per step: filter+correlation+full SVD timed in full, {detail['candidates']} of {detail['k']} candidates "
              f"(dist + Lance-Williams CONISS + broken stick + per-level CH) on {threads} threads, sweep time "
              f"scaled by k/candidates; wall for {args.steps} sampled steps {wall:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {args.bins}-bin Hi-C matrix, nested block TADs, power-law decay, max_pcs=200 "
                               "(BASELINE.json configs[1])", "bins": args.bins, "max_pcs": MAX_PCS,
                   "note": "reference is pure R and R is not installed: CPU restatement (oracle/) in the "
                           "reference's algorithmic shape, runs on rank 0 only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(device), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def h2d_bytes(n):
    """bytes tp_filter copies to the device for a row-major n x n host matrix: band copies of the upper triangle
    (csrc/filter.cu): rows [r, r + band) x columns [r, n)"""
    band = max(n // 32, 64)
    return sum((n - r) * 8 * min(band, n - r) for r in range(0, n, band))


def measure_fp64_peak(torch):
    """cuBLAS DGEMM TFLOP/s on this box (library GEMM used only as the roofline denominator)."""
    a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    b = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = max(best, 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from tadpole_b200 import ContextPool, TADpole, api
    from tadpole_b200.synth import synth_hic

    api.QUIET = True
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    S = max(1, args.streams)
    pool = ContextPool(local_rank, S)
    ctx = pool.contexts[0]
    # one host thread per call in flight waits for its stream most of the time; spinning waits (the CUDA default) need a
    # core each, so with several ranks on one host the threads would outnumber the cores
    ncpu = os.cpu_count() or 1
    sync_blocking = (world * S > ncpu // 2) if args.sync_blocking == "auto" else args.sync_blocking == "1"
    if sync_blocking:
        for c in pool.contexts:
            c.set("sync_blocking", 1)
    n = args.bins

    # synthetic inputs: POOL different matrices per rank; pinned host copies for e2e, device copies for value
    host = []
    for s in range(POOL):
        m = synth_hic(n, seed=1000 * rank + 1 + s)
        t = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        t.numpy()[:] = m
        host.append(t)
    dev = [t.cuda(non_blocking=False) for t in host]
    torch.cuda.synchronize()
    streams = [torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", local_rank)) for c in pool.contexts]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(c, i):
        d = dev[i % POOL]
        return c.call(device_ptr=d.data_ptr(), n=n, colmajor=0, max_pcs=MAX_PCS)

    def step_e2e(c, i):
        return TADpole(host[i % POOL].numpy(), max_pcs=MAX_PCS, ctx=c)

    def timed_pass(fn, steps, contexts, event_streams):
        """`steps` calls dealt out to the contexts; device time from a CUDA event recorded before the first call to the
        last end event over the contexts' streams; also wall seconds."""
        sub = ContextPool.__new__(ContextPool)
        sub.device, sub.contexts = local_rank, contexts
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in contexts]
        barrier()
        e0.record(event_streams[0])
        t0 = time.perf_counter()
        out = sub.map(fn, range(steps))
        for c, ev, st in zip(contexts, ends, event_streams):
            ev.record(st)
        for c in contexts:
            c.sync()
        wall = time.perf_counter() - t0
        barrier()
        return max(e0.elapsed_time(ev) for ev in ends), wall, out

    # ---- warm-up: every context allocates its buffers --------------------------------------------
    for w in range(max(args.warmup, 3)):
        pool.map(lambda c, i: step_dev(c, i), range(S))
    # ---- single call at a time: latency, per-kernel device times (CUDA events around each launch) ----
    lat_steps = max(5, min(args.steps, 20))
    clocks = ClockSampler(local_rank) if rank == 0 else None
    ctx.profile(1)
    l0 = sum(c.launches for c in pool.contexts)
    ms, _, outs = timed_pass(lambda c, i: step_dev(c, args.warmup + i), lat_steps, [ctx], streams[:1])
    res = outs[-1]
    prof = ctx.profile(0)
    stage = ctx.timings()
    # ---- value: S calls in flight ------------------------------------------------------------------
    ms_thr, _, _ = timed_pass(lambda c, i: step_dev(c, args.warmup + i), args.steps, pool.contexts, streams)
    launches = sum(c.launches for c in pool.contexts) - l0
    clk = clocks.stop() if clocks else None
    # ---- e2e: public API, pinned host input, copies inside the timed region -----------------------
    pool.map(lambda c, i: step_e2e(c, i), range(S))
    _, t_e2e, tps = timed_pass(lambda c, i: step_e2e(c, args.warmup + i), args.steps, pool.contexts, streams)
    _, t_e2e_single, _ = timed_pass(lambda c, i: step_e2e(c, args.warmup + i), lat_steps, [ctx], streams[:1])
    tp = tps[-1]
    d2h = int(n + tp.scores.size * 8 + (res["nf"] - 1) * 8)

    tmax = torch.tensor([ms_thr, t_e2e * 1e3, ms, t_e2e_single * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms_all, ms_single_all, e2e_single_all = (float(v) for v in tmax)
    args_steps_single = lat_steps

    if rank == 0:
        value = world * args.steps / (ms_all * 1e-3)
        e2e_val = world * args.steps / (e2e_ms_all * 1e-3)
        # roofline of the dominant kernel class (device time measured live with CUDA events around
        # every launch of the class, in the timed region above)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
        nf, k = res["nf"], res["k"]
        fp64_peak = measure_fp64_peak(torch)
        b_blk = -(-(k + max(32, k // 4)) // 32) * 32 if nf > 512 else nf      # width of the Rayleigh-Ritz problems
        nlev = float(np.count_nonzero(~np.isnan(res["scores"]))) if res.get("scores") is not None else 0.0

        def roof_of(cls):
            t_ms, cnt = prof[cls]
            if not cnt:
                return None
            per = t_ms / cnt                                   # ms per launch (CUDA events around each launch)
            src = peak_src
            if cls == "dgemm":
                # algorithmic flops summed by the library per launch (2MNK; M N (K+1) for SYRK shapes);
                # denominator: cuBLAS DGEMM measured in this same run (no FP64 peak in MEASURED_PEAKS.json)
                r = {"kernel": "dgemm_kernel (FP64 DMMA mma.sync m8n8k4)", "bound": "tensor",
                     "achieved": prof["gemm_gflop"][0] / t_ms, "peak": fp64_peak, "unit": "TFLOP/s", "traffic": None}
                src = "cuBLAS DGEMM 4096^3 via torch.matmul, best of 6, measured in this run"
            elif cls == "igemm":
                # tcgen05 int8 launches (exact Gram of the correlation, sliced operator of the PCA; the class also
                # holds their digit-slicing kernels): executed int8 operations / class time against the NOMINAL dense
                # int8 rate (no measured int8 peak in MEASURED_PEAKS.json)
                r = {"kernel": "ig_gram_kernel + io_gemm_kernel (tcgen05.mma kind::i8; digit slicing timed apart as islice)", "bound": "tensor",
                     "achieved": prof["igemm_gop"][0] / t_ms, "peak": 4500.0, "unit": "TOP/s", "traffic": None}
                src = "nominal B200 dense int8 (4.5 POP/s); executed digit-product operations, not FP64 flops"
            else:
                if cls == "coniss_sweep":      # 8 Nf k(k+1)/2 read + 8 k (Nf-1) written per sweep
                    alg, name = 8.0 * nf * k * (k + 1) / 2 + 8.0 * k * (nf - 1), "coniss_sweep_kernel"
                elif cls == "ch":              # S5: 8 (Nf-1) per candidate + 24 k per scored level
                    alg, name = 8.0 * (nf - 1) * k + 24.0 * k * nlev, "ch_kernel"
                elif cls == "rowmean":         # S1: one pass over the N x N input
                    alg, name = 8.0 * n * n, "rowmean_kernel"
                elif cls == "compact":         # S1: Nf x Nf gather written once
                    alg, name = 8.0 * nf * nf, "compact_kernel"
                elif cls == "chol":            # read G, write L and L^-1 (lower triangles): 3 * 8 * b(b+1)/2
                    alg, name = 12.0 * b_blk * (b_blk + 1), "cholinv8_kernel"
                else:                          # eigensolver: read T, write V: 2 * 8 * b^2
                    alg, name = 16.0 * b_blk * b_blk, "osj_kernel"
                r = {"kernel": name, "bound": "hbm", "achieved": alg / (per * 1e-3) / 1e9, "peak": hbm_peak,
                     "unit": "GB/s", "traffic": None}
                if cls == "coniss_sweep":
                    r["merges_per_s"] = k * (nf - 1) / (per * 1e-3)
                if cls in ("jacobi", "chol", "coniss_sweep"):
                    r["note"] = "latency-bound: serial dependent steps on L2 / shared-memory resident data"
            r["frac"] = r["achieved"] / r["peak"]
            r["peak_source"] = src
            r["share_of_step"] = t_ms / ms          # of the one-call-at-a-time pass the classes were timed in
            r["avg_launch_ms"] = per
            r["launches_per_step"] = cnt / args_steps_single
            return r

        classes = ("dgemm", "igemm", "jacobi", "chol", "coniss_sweep", "ch", "rowmean", "compact")
        roofs = {c: roof_of(c) for c in classes}
        roofs = {c: r for c, r in roofs.items() if r}
        try:        # DRAM traffic per launch measured by ncu --set full (profiles/), N = 2000 workload only
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
                traffic = json.load(fh)
            if n == 2000:
                for r in roofs.values():
                    if r["kernel"] in traffic:
                        r["traffic"] = traffic[r["kernel"]]
                        r["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
        except OSError:
            pass
        top = max(roofs, key=lambda c: roofs[c]["share_of_step"])
        roof = roofs[top]

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic {n}-bin Hi-C matrix, nested block TADs, power-law decay, max_pcs=200 "
                                   "(BASELINE.json configs[1]); one TADpole() call per step",
                       "bins": n, "good_bins": nf, "max_pcs": MAX_PCS, "n_pcs_found": res["n_pcs"],
                       "optimal_n_clusters": res["n_clusters"],
                       "l2": f"pool of {POOL} different {n}x{n} f64 matrices per rank ({POOL * n * n * 8 >> 20} MiB > L2) "
                             "cycled, no step re-reads a warm input",
                       "calls_in_flight_per_gpu": S,
                       "host_wait": ("blocking-sync event (threads sleep)" if sync_blocking else "cudaStreamSynchronize (spin)")
                                    + f", {ncpu} logical CPUs for {world * S} waiting threads",
                       "parallelism": f"{world} GPU(s) x {S} independent calls in flight (own context, stream and host thread "
                                      "each), no collective on the data path"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(n), "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_all / args.steps},
            "single_call": {"ms_per_call": ms_single_all / args_steps_single, "calls_per_s": world * args_steps_single / (ms_single_all * 1e-3),
                            "e2e_ms_per_call": e2e_single_all / args_steps_single,
                            "note": "one call at a time per GPU (latency); rooflines and kernel_ms_per_step are from this pass"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof,
            "roofline_all_kernels": roofs,
            "fp64_dgemm_peak_tflops_measured": fp64_peak,
            "stage_ms_last_step": {k_: round(v, 4) for k_, v in stage.items()},
            "kernel_ms_per_step": {c: round(v[0] / args_steps_single, 4) for c, v in prof.items() if v[1]},
            "h2d_note": "only the upper triangle of the matrix is uploaded (band copies)",
            "kernel_launches_per_step": {c: v[1] / args_steps_single for c, v in prof.items() if v[1]},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            est, detail = cpu_sample(host[0].numpy().copy(), args.cpu_sample, threads)
            line["cpu_baseline"] = {
                "value": 1.0 / est, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": (f"oracle/ (numpy + C restatement; R unavailable): filter+correlation+full SVD timed in full "
                           f"({detail['front_s']} s), {detail['candidates']} of {detail['k']} candidates on {threads} "
                           f"threads ({detail['sweep_sample_s']} s) scaled by k/candidates")}
            # the tensor-core kernels at a size where they are the step (the 2000-bin workload leaves them 64-CTA grids)
            if args.large_n > 0:
                nl = args.large_n
                ml = synth_hic(nl, seed=77)
                c0 = pool.contexts[0]
                for _ in range(2):
                    rl = c0.call(ml, max_pcs=MAX_PCS)
                c0.profile(1)
                rl = c0.call(ml, max_pcs=MAX_PCS)
                pl = c0.profile(0)
                stl = c0.timings()
                big = {"bins": nl, "good_bins": rl["nf"], "n_pcs_found": rl["n_pcs"],
                       "stage_ms": {k_: round(v, 3) for k_, v in stl.items()},
                       "kernel_ms": {c: round(v[0], 3) for c, v in pl.items() if v[1]},
                       "kernel_launches": {c: v[1] for c, v in pl.items() if v[1]}}
                if pl["igemm"][1]:
                    a = pl["igemm_gop"][0] / pl["igemm"][0]
                    big["igemm_roofline"] = {"kernel": "ig_gram_kernel + io_gemm_kernel<5|8> (tcgen05.mma kind::i8)", "bound": "tensor",
                                             "achieved": a, "peak": 4500.0, "unit": "TOP/s", "frac": a / 4500.0,
                                             "peak_source": "nominal B200 dense int8; executed digit-product operations over the "
                                                            "CUDA-event time of the tcgen05 launches of one call"}
                if pl["dgemm"][1]:
                    a = pl["gemm_gflop"][0] / pl["dgemm"][0]
                    big["dgemm_roofline"] = {"kernel": "dgemm_kernel (FP64 DMMA)", "bound": "tensor", "achieved": a, "peak": fp64_peak,
                                             "unit": "TFLOP/s", "frac": a / fp64_peak,
                                             "peak_source": "cuBLAS DGEMM 4096^3 measured in this run"}
                try:      # DRAM traffic per launch of the tcgen05 kernels from the ncu --set full capture at this size
                    with open(os.path.join(ROOT, "profiles", f"r01_traffic_n{nl}.json")) as fh:
                        big["traffic_bytes_per_launch_ncu"] = {k_: v for k_, v in json.load(fh).items() if not k_.startswith("_")}
                except OSError:
                    pass
                line["large_n_call"] = big
                del ml
            # input side: the same call starting from the matrix FILE, text parsed on the GPU
            try:
                import tempfile
                hm = host[0].numpy()
                text = "\n".join("\t".join(map(str, row)) for row in hm.astype(np.int64).tolist()) + "\n"
                with tempfile.NamedTemporaryFile("w", suffix=".tsv", delete=False) as fh:
                    fh.write(text)
                c0 = pool.contexts[0]
                t_file = []
                for _ in range(6):
                    t0 = time.perf_counter()
                    ptr, n_in = c0.ingest_tsv(fh.name)
                    c0.call(device_ptr=ptr, n=n_in, colmajor=0, max_pcs=MAX_PCS)
                    t_file.append((time.perf_counter() - t0) * 1e3)
                ist = c0.ingest_stats()
                os.unlink(fh.name)
                alg = ist["text_bytes"] + 8.0 * n * n            # file bytes read once + the FP64 matrix written once
                line["from_file"] = {
                    "ms_per_call": float(np.median(t_file[1:])), "text_bytes": ist["text_bytes"],
                    "ingest_wall_ms": ist["wall_ms"], "parse_kernels_ms": ist["parse_ms"], "host_converted_fields": ist["host_fields"],
                    "roofline": {"kernel": "newline_* + parse_rows_kernel (ingest.cu)", "bound": "hbm",
                                 "achieved": alg / (ist["parse_ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg / (ist["parse_ms"] * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                                 "note": "event span of 4 launches incl. one host read-back of the row count"},
                    "note": "wall clock from the open() of the TSV file to the returned result, page cache warm"}
            except OSError as e:
                line["from_file"] = {"unavailable": str(e)}
        print(json.dumps(line), flush=True)
    pool.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
