"""DRAM traffic of the generator launches of one bench step, from an ncu launch list
(--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum).  The generator's launches are the ones
after the acoustic model's last kernel (the postnet's final conv): from conv_pre (the first tc_conv launch after the
last EPI=1 launch) to conv_post_kernel.  Writes profiles/r2_traffic.json (read by bench.py for roofline.traffic)."""
import csv
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
path = Path(sys.argv[1])
pairs = sys.argv[2] if len(sys.argv) > 2 else "auto"
rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
hdr = next(r for r in csv.reader(open(path)) if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(hdr)}
launch = {}
for r in rows:
    d = launch.setdefault(int(r[ix["ID"]]), dict(name=r[ix["Kernel Name"]]))
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    d[r[ix["Metric Name"]]] = v * mul
ids = sorted(launch)
# the last full step: find the last conv_post, walk back to the launch after the last acoustic (EPI = 1) tensor-core conv before it
last_post = max(i for i in ids if "conv_post" in launch[i]["name"])
seg = [i for i in ids if i <= last_post]
def epi(name):
    return int(name.split("tc_conv_kernel<")[1].split(">")[0].split(",")[1]) if "tc_conv_kernel<" in name else -1


start = max(i for i in seg if epi(launch[i]["name"]) == 1) + 1
gen = [i for i in seg if i >= start]
byts = sum(launch[i].get("dram__bytes_read.sum", 0) + launch[i].get("dram__bytes_write.sum", 0) for i in gen)
ns = sum(launch[i].get("gpu__time_duration.sum", 0) for i in gen)
out = dict(workload=dict(batch=32, mel_frames=312, precision="bf16x3", pairs=pairs), generator_launches=len(gen), generator_dram_bytes_per_step=byts,
           generator_ns_serialised=ns, kernels=sorted(set(launch[i]["name"].split("(")[0] for i in gen)),
           source=f"profiles/{path.name} (ncu dram__bytes_read.sum + dram__bytes_write.sum over the generator launches of one step: conv_pre .. conv_post)")
(REPO / "profiles" / "r2_traffic.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
