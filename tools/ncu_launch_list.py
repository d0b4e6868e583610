"""Compact an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`):
per-kernel totals to stdout, every launch (id, kernel, grid, block, duration) to a gzipped CSV.
python tools/ncu_launch_list.py X.csv out.csv.gz "<cmd that was profiled>" [last_step]
last_step: keep only the launches of the last train step (steps are delimited by their two leading MDCT launches)."""
import collections
import csv
import gzip
import re
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.DictReader(l for l in open(src) if not l.startswith("==")))
if len(sys.argv) > 4 and sys.argv[4] == "last_step":
    names = [r["Kernel Name"] for r in rows]
    fwd = [i for i, n in enumerate(names) if "mdct4_fwd_kernel" in n]
    starts = [i for k, i in enumerate(fwd) if k == 0 or fwd[k - 1] != i - 1]
    rows = rows[starts[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0])
with gzip.open(dst, "wt") as f:
    f.write("id,kernel,grid,block,duration_us\n")
    for r in rows:
        n = re.sub(r"^(mdctk|umma|nnk|trk)::", "", re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""))
        us = float(r["Metric Value"]) / 1000.0
        agg[n][0] += 1
        agg[n][1] += us
        f.write(f'{r["ID"]},"{n}","{r["Grid Size"]}","{r["Block Size"]}",{us:.2f}\n')
tot = sum(v[1] for v in agg.values())
print(f"# ncu launch list of `{cmd}`: {len(rows)} launches, {tot / 1000:.2f} ms of kernel time (serialised, cold caches, under ncu)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f} % {v[0]:6d} x  {k[:110]}")
