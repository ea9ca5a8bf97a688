"""Top stall lines from `ncu --page source --csv` output (SASS view): python tools/ncu_src_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:N]
for r in top:
    s = int(r[ix["# Samples"]] or 0)
    st = {h[6:]: int(r[ix[h]] or 0) for h in stall_cols if int(r[ix[h]] or 0)}
    print(f"{s:6d} {100.0*s/tot:5.1f}%  {r[ix['Address']][-5:]}  {r[ix['Source']][:70]:70s} {st}")
