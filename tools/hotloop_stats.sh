#!/bin/bash
# development aid: spill / instruction statistics of the TC=10 ADMM kernel's plain-iteration loop and sweep step
cd /root/repo
cuobjdump -sass qpcontrol.jl_b200/csrc/libqpcontrol_b200.so | awk '/Function : /{f=($3 ~ /admm_reg_kernelILi10ELi8E/)} f' | grep -v "^\s*/\* 0x" | sed 's#/\*[0-9a-f]*\*/##; s#/\* 0x[0-9a-f]* \*/##' > /tmp/t10.sass
python3 - <<'PY'
import re
L=open('/tmp/t10.sass').read().split('\n')
bars=[i for i,l in enumerate(L) if 'BAR.SYNC' in l]
segs=list(zip([0]+bars, bars+[len(L)]))
def stats(a,b):
    s=L[a:b]
    return dict(n=len(s), dfma=sum('DFMA' in l for l in s), dmul=sum('DMUL' in l for l in s), ldl=sum('LDL' in l for l in s), stl=sum('STL' in l for l in s), lds=sum('LDS' in l for l in s), sts=sum('STS' in l for l in s), shfl=sum('SHFL' in l for l in s))
for a,b in segs:
    st=stats(a,b)
    if st['dfma']>=30: print(a,b,st)
PY
