import csv,subprocess,sys,collections
rep,rx=sys.argv[1],sys.argv[2]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+rx],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
hdr=rows[hi]; idx={h:i for i,h in enumerate(hdr)}
sass=[]
for r in rows[hi+1:]:
    if r and r[0]=="Kernel Name": break
    if r and r[0].startswith("0x"): sass.append(r)
base=int(sass[0][0],16)
tot=sum(int(r[idx['# Samples']]) for r in sass); ti=sum(int(r[idx['Instructions Executed']]) for r in sass)
ph=0; agg=collections.OrderedDict()
for r in sass:
    a=int(r[0],16)-base
    op=r[1].strip(); m=op.split()[1] if op.startswith('@') else op.split()[0]
    d=agg.setdefault(ph,dict(s=0,i=0,ops=collections.Counter(),start=a,st=collections.Counter()))
    s=int(r[idx['# Samples']]); ie=int(r[idx['Instructions Executed']])
    d['s']+=s; d['i']+=ie; d['ops'][m.split('.')[0]]+=ie; d['end']=a
    for h in hdr:
        if h.startswith('stall_') and 'Not Issued' not in h and int(r[idx[h]])>0: d['st'][h[6:]]+=int(r[idx[h]])
    if m.startswith('BAR'): ph+=1
    if m.startswith('EXIT'): break
for p,d in agg.items():
    print(f"phase{p} [{d['start']:#x}-{d['end']:#x}] samp {d['s']/tot:.1%} inst {d['i']/ti:.1%} ({d['i']/8736/4:.0f}/warp) | "+', '.join(f"{k}:{v/max(d['i'],1):.0%}" for k,v in d['ops'].most_common(7))+" | "+', '.join(f"{k}:{v/max(d['s'],1):.0%}" for k,v in d['st'].most_common(4)))
