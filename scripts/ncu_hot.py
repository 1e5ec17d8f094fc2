"""Top SASS instructions by stall samples from `ncu --page source --csv`."""
import csv, sys
path = sys.argv[1]
with open(path) as f:
    r = csv.reader(f)
    rows = list(r)
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(d[ix["# Samples"]] or 0) for d in data)
print("total samples", tot)
top = sorted(data, key=lambda d: -int(d[ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for d in top:
    n = int(d[ix["# Samples"]] or 0)
    st = sorted(((int(d[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{100*n/tot:5.1f}%  {d[ix['Source']].strip()[:70]:70s} exec={d[ix['Instructions Executed']]:>8s} " + " ".join(f"{c}={v}" for v, c in st if v))
