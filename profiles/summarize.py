"""Turn ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
  python profiles/summarize.py full gpurun_out/prof_x.ncu-rep ... > profiles/rNN_full.md

`launches` aggregates the `--metrics gpu__time_duration.sum` launch list per kernel (count, total,
share: cold-cache serialised times, so only the SHARES are comparable with bench.py).
`full` extracts the counters the roofline discussion uses from `--set full` reports."""
import csv
import subprocess
import sys
from collections import OrderedDict

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("sm__inst_executed_pipe_tensor_op_dmma.sum", "DMMA instructions"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of elapsed"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem -> tensor core wavefronts % of peak"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor memory active % of elapsed"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def simplify(name):
    name = name.replace("void ", "")
    p = name.find("(")
    return name[:p] if p > 0 else name


def launches(path):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        rows.append((simplify(r["Kernel Name"]), r["Grid Size"], r["Block Size"], v))
    agg = OrderedDict()
    for k, g, b, v in rows:
        a = agg.setdefault(k, [0, 0.0, g, b])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"launches: {len(rows)}   total device time (serialised, cold): {tot / 1e3:.3f} ms\n")
    print("| kernel | launches | total us | share | avg us | grid (first) | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0]:.1f} | {a[2]} | {a[3]} |")


def full(paths, dedupe=False):
    """paths: .ncu-rep reports, or the `ncu -i ... --page raw --csv` text already made on the GPU box (*.csv; gpurun brings
    back at most 64 MiB, a --set full report of a whole call is larger).  dedupe: one table per distinct kernel and launch
    shape (the launch with the median duration) plus the launch count and the duration range."""
    for p in paths:
        if p.endswith(".csv"):
            with open(p) as fh:
                out = fh.read()
        else:
            out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
        if len(rows) < 3:
            print(f"## {p}: no data\n")
            continue
        hdr, units = rows[0], rows[1]
        print(f"## {p}\n")
        body = rows[2:]
        note = {}
        if dedupe:
            di = hdr.index("gpu__time_duration.sum")
            groups = OrderedDict()
            for r in body:
                groups.setdefault((r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]), []).append(r)
            body = []
            for key, rs in groups.items():
                rs = sorted(rs, key=lambda r: float(r[di].replace(",", "")))
                pick = rs[len(rs) // 2]
                body.append(pick)
                note[id(pick)] = f"{len(rs)} launches of this shape in the call, duration {rs[0][di]} .. {rs[-1][di]} {units[di]}; the median one:"
        for r in body:
            name = r[hdr.index("Kernel Name")]
            print(f"### `{simplify(name)}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            if id(r) in note:
                print(note[id(r)] + "\n")
            print("| counter | value |")
            print("|---|---|")
            for key, label in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    print(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
            print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "full1":
        full(sys.argv[2:], dedupe=True)
    else:
        full(sys.argv[2:])
