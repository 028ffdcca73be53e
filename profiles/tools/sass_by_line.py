#!/usr/bin/env python
"""Join an `ncu --page source --csv` (SASS view) dump with `nvdisasm -g` line info and aggregate the
per-instruction counters by CUDA source line.  Usage:
    cuobjdump -xelf all libglomecuda.so ; nvdisasm -g -c glome_cuda.sm_100a.cubin > dis.txt
    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<k> > src.csv
    sass_by_line.py dis.txt src.csv <mangled-kernel-substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict

dis, src, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# --- nvdisasm: instruction index -> (file, line) of the innermost location ---
lines = open(dis, errors="replace").read().splitlines()
start = None
for i, l in enumerate(lines):
    if l.startswith("//---") and ".text." in l and kname in l:
        start = i
        break
assert start is not None, "kernel not found in disassembly"
loc = ("?", 0)
insn_loc = []
for l in lines[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l):
        insn_loc.append(loc)

# --- ncu SASS rows for the first kernel instance in the csv whose name matches ---
rows = list(csv.reader(open(src)))
hdr = None
data = []
take = False
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        if take and data:
            break
        take = True
        continue
    if r[0] == "Address":
        hdr = r
        continue
    if take and hdr and len(r) >= len(hdr) - 2:
        data.append(r)
col = {n: i for i, n in enumerate(hdr)}
n = min(len(data), len(insn_loc))
agg = defaultdict(lambda: [0, 0, 0, 0, 0])
tot_i = tot_t = tot_s = 0
def num(x):
    try:
        return float(x)
    except Exception:
        return 0.0
for k in range(n):
    r = data[k]
    ie, te = num(r[col["Instructions Executed"]]), num(r[col["Thread Instructions Executed"]])
    ss = num(r[col["# Samples"]])
    lsb = num(r[col["stall_long_sb"]]) if "stall_long_sb" in col else 0
    a = agg[insn_loc[k]]
    a[0] += ie; a[1] += te; a[2] += ss; a[3] += lsb; a[4] += 1
    tot_i += ie; tot_t += te; tot_s += ss
print("kernel %s: %d SASS instructions (%d with line info), warp-instr %.3g, avg active threads %.2f"
      % (kname, len(data), len(insn_loc), tot_i, tot_t / max(tot_i, 1)))
print("%-28s %8s %7s %7s %8s %8s %6s" % ("file:line", "winstr%", "active", "sass", "samples%", "long_sb%", ""))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%-28s %7.2f%% %7.2f %7d %7.2f%% %7.2f%%" % ("%s:%d" % (f, ln), 100 * a[0] / max(tot_i, 1), a[1] / max(a[0], 1), a[4],
                                                 100 * a[2] / max(tot_s, 1), 100 * a[3] / max(tot_s, 1)))
