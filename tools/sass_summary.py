"""SASS opcode summary per kernel of the shipped library: the Blackwell-specific mnemonics that prove tcgen05 / TMEM / TMA.

    python tools/sass_summary.py [rangeclip_b200/librangeclip_b200.so] > profiles/r2_sass_opcodes.txt

UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG / UTMASTG /
UTMAREDG = cp.async.bulk.tensor load / store / reduce, SYNCS = mbarrier ops, UTCATOMSWS/UTCCP... as they appear."""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "rangeclip_b200", "librangeclip_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMACCTL", "UTMACMDFLUSH", "SYNCS", "UBLKCP",
         "HMMA", "MUFU", "REDG", "ATOMG", "LDGSTS", "STSM", "LDSM", "USETMAXREG", "FHFMA", "FHADD", "HFMA2", "STL", "LDL"]
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        total[kern] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or (w in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "MUFU", "HMMA") and op.startswith(w)):
                counts[kern][w] += 1
                break
print(f"# cuobjdump -sass {os.path.basename(so)} (sm_100a); instruction counts per kernel (static SASS, not executed counts)")
print(f"{'kernel':78s} {'sass':>6s}  " + " ".join(f"{w:>8s}" for w in WATCH if any(c[w] for c in counts.values())))
cols = [w for w in WATCH if any(c[w] for c in counts.values())]
for k, c in counts.items():
    name = demangle(k)
    if name.endswith(")"):            # drop the trailing parameter list, keep the template arguments
        depth = 0
        for i in range(len(name) - 1, -1, -1):
            depth += name[i] == ")"
            depth -= name[i] == "("
            if depth == 0:
                name = name[:i]
                break
    name = name.replace("void ", "").replace("rc::", "").replace("(bool)", "").replace("(int)", "")
    print(f"{name[:78]:78s} {total[k]:6d}  " + " ".join(f"{c[w]:8d}" for w in cols))
