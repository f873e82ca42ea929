#!/usr/bin/env python
"""Summarises `ncu --page source --csv` of a kernel: total stall reasons, and sample share per SASS window (cut at
WARPSYNC / backward branches), with the dominant stall of each window.   python tools/ncu_warp_stalls.py rep [n_windows]"""
import csv, sys, subprocess, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed"); iA = hdr.index("Address"); iT = hdr.index("Source")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = []
for r in rows[2:]:
    if len(r) <= iE: continue
    ins.append(dict(addr=r[iA], text=r[iT].strip(), samples=int(r[iS] or 0), execd=int(r[iE] or 0),
                    stalls={h: int(r[i] or 0) for i, h in stall_cols}))
tot = sum(x["samples"] for x in ins)
print("instructions", len(ins), "samples", tot)
agg = {}
for x in ins:
    for h, v in x["stalls"].items(): agg[h] = agg.get(h, 0) + v
print("stall totals:", ", ".join(f"{h[6:]} {100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n = len(ins); step = (n + W - 1) // W
print(f"{'window':>14s} {'share':>6s} {'exec/instr':>10s}  mix / top stalls")
for a in range(0, n, step):
    w = ins[a:a + step]
    s = sum(x["samples"] for x in w)
    ex = sum(x["execd"] for x in w) / max(1, len(w))
    st = {}
    for x in w:
        for h, v in x["stalls"].items(): st[h] = st.get(h, 0) + v
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    txt = [x["text"] for x in w]
    mix = f"dfma {sum('DFMA' in t for t in txt)} dmul {sum('DMUL' in t for t in txt)} lds {sum('LDS' in t for t in txt)} sts {sum('STS' in t for t in txt)} ldl {sum('LDL' in t for t in txt)} stl {sum('STL' in t for t in txt)} ldg {sum('LDG' in t for t in txt)} shfl {sum('SHFL' in t for t in txt)} mufu {sum('MUFU' in t for t in txt)} sync {sum('WARPSYNC' in t for t in txt)}"
    print(f"{a:6d}-{a+len(w):6d} {100*s/tot:5.1f}% {ex:10.0f}  {mix} | " + ", ".join(f"{h[6:]} {100*v/max(1,s):.0f}%" for h, v in top))
