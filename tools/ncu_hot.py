"""Top stall sites of an `ncu --page source --csv` dump:  ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {name: i for i, name in enumerate(hdr)}
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
body = rows[2:]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print("total samples", tot)
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    top = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[col['# Samples']]):6d} exec={r[col['Instructions Executed']]:>8s} {r[col['Source']].strip()[:90]:90s} {top}")
