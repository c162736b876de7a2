"""Developer tool: per-SASS-instruction execution counts of one kernel from an .ncu-rep (ncu --page source)."""
import csv, subprocess, sys
from collections import Counter
rep, kre = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else None
hi = int(sys.argv[4]) if len(sys.argv) > 4 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 10 and r[0].startswith("0x")]
# several launches of the kernel may be in the report: keep the first
seen, first = set(), []
for r in data:
    if r[0] in seen: break
    seen.add(r[0]); first.append(r)
data = first
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[iE]) for r in data)
print("total warp instr", tot, "sass lines", len(data))
if lo is None:
    c = Counter()
    for r in data:
        op = [o for o in r[iS].strip().split() if not o.startswith("@")][0].split(".")[0]
        c[op] += int(r[iE])
    for k, v in c.most_common(24): print(f"{k:12s} {v:>12d} {100*v/tot:5.1f}%")
    cur, start = None, 0
    for i, r in enumerate(data + [None]):
        e = int(r[iE]) if r else -1
        if e != cur:
            if cur is not None and cur * (i - start) > 0.01 * tot: print(f"sass {start:4d}-{i-1:4d} n={i-start:4d} exec={cur} ({100*cur*(i-start)/tot:.1f}%)")
            cur, start = e, i
else:
    for i in range(lo, min(hi, len(data))): print(i, data[i][iE], data[i][iS].strip())
