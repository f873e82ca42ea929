import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]
iS=hdr.index('# Samples'); iE=hdr.index('Instructions Executed')
ins=[(r[1].strip(),int(r[iS]),int(r[iE])) for r in rows[2:] if len(r)>iE]
tot=sum(s for _,s,_ in ins)
segs=[];a=0
for i,(t,s,e) in enumerate(ins):
    if 'BAR.SYNC' in t:
        segs.append((a,i+1)); a=i+1
segs.append((a,len(ins)))
print('total samples',tot,'n instr',len(ins))
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.02
for a,b in segs:
    s=sum(x[1] for x in ins[a:b])
    if s/tot>thr:
        ex=max(x[2] for x in ins[a:b])
        txt=[x[0] for x in ins[a:b]]
        print(f"{a:6d}-{b:6d} {100*s/tot:5.1f}% maxexec={ex} n={b-a} dfma={sum('DFMA' in t for t in txt)} lds={sum('LDS' in t for t in txt)} ld={sum(t.startswith('LD.') or ' LD.E' in t or t.startswith('@P') and ' LD.' in t for t in txt)} st={sum(' ST.' in t or t.startswith('ST.') for t in txt)}")
