#!/usr/bin/env python
"""Executed warp-instructions and stall samples of a kernel by SOURCE LINE (ncu --page source --csv joined with the cubin's
line table), aggregated per file:line and per function-ish region.  python tools/ncu_lines_exec.py report.ncu-rep <function substring> [top] [lib.so]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, fn = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "qpcontrol.jl_b200", "csrc", "libqpcontrol_b200.so")
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
iS, iA, iE = hdr.index("# Samples"), hdr.index("Address"), hdr.index("Instructions Executed")
ex = collections.Counter(); sm = collections.Counter(); nstatic = collections.Counter()
base = None; tot_e = tot_s = 0
for r in rows[2:]:
    if len(r) <= iE or not r[iA].startswith("0x"): continue
    a = int(r[iA], 16)
    if base is None: base = a
    ln = addr2line.get(a - base, ("?", 0))
    e, s = int(r[iE] or 0), int(r[iS] or 0)
    ex[ln] += e; sm[ln] += s; nstatic[ln] += 1; tot_e += e; tot_s += s
print("static instructions", sum(nstatic.values()), "executed warp-instr", tot_e, "samples", tot_s)
src = {}
def text(f, ln):
    if f not in src:
        p = os.path.join(ROOT, "qpcontrol.jl_b200", "csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
print("--- by executed instructions")
for (f, ln), e in ex.most_common(top):
    print(f"{100*e/tot_e:5.1f}% exec {100*sm[(f,ln)]/max(1,tot_s):5.1f}% samples {nstatic[(f,ln)]:5d} static  {f}:{ln:<4d} {text(f, ln)}")
