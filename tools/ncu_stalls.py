"""Aggregate `ncu --page source --csv` (SASS view, warp-stall sampling) of one kernel by code region.

    ncu -i prof.ncu-rep --page source --csv > src.csv ;  python tools/ncu_stalls.py src.csv [top_n]

Regions are split at SASS markers: USETMAXREG (register re-allocation = start of a role's code) in program order.
Prints per region: samples, executed warp-instructions, top stall reasons, and the hottest instructions."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
regions, cur = [], dict(name="entry", rows=[])
for r in data:
    sass = r[col["Source"]].strip()
    if "USETMAXREG" in sass:
        regions.append(cur)
        cur = dict(name=sass.split(";")[0][:40], rows=[])
    cur["rows"].append(r)
regions.append(cur)
total = sum(int(r[col["# Samples"]] or 0) for r in data)
print(f"# total samples {total}")
for reg in regions:
    n = sum(int(r[col["# Samples"]] or 0) for r in reg["rows"])
    ins = sum(int(r[col["Instructions Executed"]] or 0) for r in reg["rows"])
    st = Counter()
    for r in reg["rows"]:
        for s in stall_cols:
            st[s] += int(r[col[s]] or 0)
    print(f"== region after [{reg['name']}]  sass {len(reg['rows'])}  samples {n} ({100.0 * n / max(total, 1):.1f}%)  warp-instr {ins}")
    print("   stalls: " + ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(7)))
    hot = sorted(reg["rows"], key=lambda r: -int(r[col["# Samples"]] or 0))[:top_n]
    for r in hot:
        rs = Counter({s: int(r[col[s]] or 0) for s in stall_cols})
        print(f"   {int(r[col['# Samples']]):7d}  x{int(r[col['Instructions Executed']] or 0):9d}  {r[col['Source']].strip()[:90]:90s} " +
              ",".join(f"{k[6:]}={v}" for k, v in rs.most_common(2)))
