"""Summarise an ncu launch list (time [+ dram bytes] per launch) for the last full bench step."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
verbose = len(sys.argv) > 2
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
# group metrics per launch ID
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r["ID"], dict(name=r["Kernel Name"], grid=r["Grid Size"]))
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m.startswith("gpu__time"):
        d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
    elif "dram__bytes" in m:
        d[m.split(".")[0]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
L = list(launch.values())
idx = [i for i, d in enumerate(L) if "embed_kernel" in d["name"]]
seg = L[idx[-2]:idx[-1]] if len(idx) > 1 else L
tot = collections.defaultdict(lambda: [0.0, 0, 0.0])
T = sum(d["ns"] for d in seg)
for d in seg:
    key = d["name"].split("(TcLaunch")[0].split("(ConvLaunch")[0].split("(const")[0].replace("<unnamed>::", "").replace("void ", "").replace("(int)", "")
    t = tot[key]
    t[0] += d["ns"]
    t[1] += 1
    t[2] += d.get("dram__bytes_read", 0) + d.get("dram__bytes_write", 0)
    if verbose:
        print(f"{d['ns']/1e3:9.1f} us {(d.get('dram__bytes_read',0)+d.get('dram__bytes_write',0))/1e6:9.1f} MB  {key} {d['grid']}")
print(f"one step: {T/1e6:.3f} ms over {len(seg)} launches (ncu-serialised, cold cache)")
for k, (ns, n, by) in sorted(tot.items(), key=lambda x: -x[1][0]):
    extra = f"  dram {by/1e9:6.2f} GB -> {by/ns:6.0f} GB/s" if by else ""
    print(f"{ns/1e6:9.3f} ms {100*ns/T:6.2f}% n={n:3d} {k}{extra}")
