"""Per-source-line stall samples of one kernel from an ncu report (needs --import-source on and -lineinfo):
   python profiles/hotlines.py report.ncu-rep kernel_regex source_file_suffix [top]"""
import collections
import csv
import subprocess
import sys

rep, kern, suffix = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg, src_of = None, None, collections.OrderedDict(), {}
launch = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or cur is None or not cur.endswith(suffix):
        continue
    if r[0].isdigit():
        off = len(hdr) - hdr.index("Warp Stall Sampling (All Samples)")
        try:
            v = float(r[len(r) - off])
        except (ValueError, IndexError):
            v = 0.0
        ln = int(r[0])
        agg[ln] = agg.get(ln, 0.0) + v
tot = sum(agg.values()) or 1.0
try:
    src = open(cur if False else [p for p in [suffix, "/root/repo/tadpole_b200/csrc/" + suffix] if __import__("os").path.exists(p)][0]).read().split("\n")
except IndexError:
    src = []
print(f"kernel {kern}: {tot:.0f} stall samples attributed to {suffix} (all captured launches)")
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    text = src[ln - 1].strip()[:110] if 0 < ln <= len(src) else ""
    print(f"{ln:5d} {v:8.0f} {100 * v / tot:5.1f}%  {text}")
