"""Top stalled SASS instructions of one kernel from `ncu -i rep --page source --csv` output.
    ncu -i prof.ncu-rep --page source --csv --kernel-name-base demangled -k regex:NAME > src.csv
    python tools/ncu_source_top.py src.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
num = lambda s: int(s) if s.strip().lstrip("-").isdigit() else 0
tot = sum(num(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(num(r[ix[h]]) for r in data) for h in stall_cols}
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:n]:
    st = {h: num(r[ix[h]]) for h in stall_cols}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(8), r[ix["Source"]].strip()[:80].ljust(80), main)
