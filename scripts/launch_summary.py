"""Summarise an ncu launch list (gpu__time_duration.sum per launch) for the last full step."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
idx = [i for i, n in enumerate(names) if "embed_kernel" in n]
seg = rows[idx[-2]:idx[-1]] if len(idx) > 1 else rows
tot = collections.defaultdict(float)
cnt = collections.Counter()
T = 0.0
for r in seg:
    v = float(r["Metric Value"].replace(",", ""))
    key = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "") + " grid=" + r["Grid Size"]
    tot[key] += v
    cnt[key] += 1
    T += v
print(f"one step: {T/1e6:.3f} ms over {len(seg)} launches (ncu-serialised, cold cache)")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:30]:
    print(f"{v/1e6:9.3f} ms {100*v/T:6.2f}% n={cnt[k]:3d} {k}")
