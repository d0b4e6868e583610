"""Per-kernel summary of the LAST step in an ncu launch list (gpu__time_duration.sum csv of tools/train_step_once.py)."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in rows]
adam = [i for i, n in enumerate(names) if "adam_flat" in n]
start, end = adam[-3] + 1, adam[-1] + 1
while end < len(rows) and "pack_weights" in names[end]:
    end += 1
agg = collections.defaultdict(lambda: [0, 0.0])
for x in rows[start:end]:
    n = re.sub(r"^void ", "", re.sub(r"\(.*", "", x["Kernel Name"]))
    agg[n][0] += 1
    agg[n][1] += float(x["Metric Value"]) / 1000.0
tot = sum(v[1] for v in agg.values())
print(f"# kernels in the last step: {end - start}, sum of gpu__time_duration: {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f} % {v[0]:4d} x  {k[:120]}")
