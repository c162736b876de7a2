"""Developer tool: hottest SASS lines (stall samples) of one kernel from an ncu report.
usage: python tools/ncu_hot.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[start]
si, ie = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
data = []
for r in rows[start + 1:]:
    if len(r) <= max(si, ie) or r[0] == "Address" or not r[si].isdigit():
        if r and r[0] in ("Kernel Name",): break
        continue
    data.append(r)
tot = sum(int(r[si]) for r in data)
print("samples", tot, "warp-insts", sum(int(r[ie] or 0) for r in data), "sass lines", len(data))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:top]
for i in sorted(idx):
    r = data[i]
    print(f"{i:5d} {int(r[si]):6d} {100.0*int(r[si])/max(tot,1):5.1f}% {r[ie]:>9} {r[1][:100]}")
