"""Per-kernel summary of the LAST train step in an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv` of
tools/train_step_once.py): steps are delimited by their two leading MDCT launches.  python tools/launch_summary2.py launches.csv [top]"""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in rows]
fwd = [i for i, n in enumerate(names) if "mdct4_fwd_kernel" in n]
starts = [i for k, i in enumerate(fwd) if k == 0 or fwd[k - 1] != i - 1]          # first of each (lr, hr) pair
start, end = starts[-1], len(rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for x in rows[start:end]:
    n = re.sub(r"^void ", "", re.sub(r"\(.*", "", x["Kernel Name"]))
    n = re.sub(r"^(mdctk|umma|nnk|trk)::", "", n)
    agg[n][0] += 1
    agg[n][1] += float(x["Metric Value"]) / 1000.0
tot = sum(v[1] for v in agg.values())
print(f"# kernels in the last step: {end - start}, sum of gpu__time_duration (serialised, cold caches under ncu): {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f} % {v[0]:4d} x  {k[:110]}")
