"""Developer tool: static SASS of one kernel from a .o / .so (cuobjdump), with backward-branch loop sizes.
usage: sass_loops.py <object> <kernel-name-substring> [print]"""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    if pat not in name: continue
    ins = []
    for ln in blk.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(f"== {name[:110]}  ({len(ins)} instructions)")
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.[A-Z.]+)?\s+(?:[!A-Z0-9, ]*?)?`?\(?\.?L?_?x?_?\d*\)?\s*(0x[0-9a-f]+)", t)
        if "BRA" in t:
            m2 = re.search(r"(0x[0-9a-f]+)\s*$", t)
            if m2:
                tgt = int(m2.group(1), 16)
                if tgt in addr_idx and addr_idx[tgt] <= i:
                    print(f"   loop: sass {addr_idx[tgt]:4d}..{i:4d}  = {i - addr_idx[tgt] + 1} instructions")
    if len(sys.argv) > 3:
        lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
        hi = int(sys.argv[5]) if len(sys.argv) > 5 else len(ins)
        for i in range(lo, min(hi, len(ins))): print(i, ins[i][1])
