#!/usr/bin/env python
"""Warp-stall samples of a kernel by SOURCE LINE: joins `ncu --page source --csv` (per SASS address) with the line table of
the cubin (`nvdisasm --print-line-info`).   python tools/ncu_lines.py report.ncu-rep <function substring> [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, fn = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "qpcontrol.jl_b200", "csrc", "libqpcontrol_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
addr2line = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn, cur = False, None
    for l in txt.split("\n"):
        if l.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", l):
            infn = fn in l
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2)))
        m = re.match(r"^\s+/\*([0-9a-f]{4,6})\*/", l)
        if m and cur: addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iA = hdr.index("# Samples"), hdr.index("Address")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter(); st = collections.defaultdict(collections.Counter); tot = 0
base = None
for r in rows[2:]:
    if len(r) <= iS: continue
    a = int(r[iA], 16) if r[iA].startswith("0x") or re.match(r"^[0-9a-f]+$", r[iA]) else None
    if a is None: continue
    if base is None: base = a
    s = int(r[iS] or 0); tot += s
    ln = addr2line.get(a - base, ("?", 0))
    agg[ln] += s
    for i, h in stall_cols: st[ln][h[6:]] += int(r[i] or 0)
src = {}
print("total samples", tot)
for (f, ln), s in agg.most_common(top):
    if f not in src:
        p = os.path.join(ROOT, "qpcontrol.jl_b200", "csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    ts = ", ".join(f"{k} {100*v/max(1,s):.0f}%" for k, v in st[(f, ln)].most_common(2))
    print(f"{100*s/tot:5.1f}%  {f}:{ln:<4d} {text}   [{ts}]")
