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
# segment at RET / function boundaries: detect by 'RET' instruction
seg=0; segs=collections.OrderedDict()
for r in sass:
    a=int(r[0],16)-base
    op=r[1].strip()
    d=segs.setdefault(seg,dict(start=a,n=0,samp=0,inst=0,dfma=0,f64=0,lds=0,ldg=0,ops=collections.Counter()))
    d['n']+=1; d['samp']+=int(r[idx['# Samples']]); ie=int(r[idx['Instructions Executed']]); d['inst']+=ie
    m=op.split()[1] if op.startswith('@') else op.split()[0]
    m0=m.split('.')[0]
    d['ops'][m0]+=ie
    d['end']=a
    if m0=='RET' or m0=='EXIT': seg+=1
tot=sum(d['samp'] for d in segs.values()); ti=sum(d['inst'] for d in segs.values())
for s,d in segs.items():
    if d['inst']==0 and d['samp']==0: continue
    top=', '.join(f"{k}:{v/ max(d['inst'],1):.0%}" for k,v in d['ops'].most_common(8))
    print(f"seg{s} [{d['start']:#x}-{d['end']:#x}] n={d['n']} samp={d['samp']/tot:.1%} inst={d['inst']/ti:.1%} | {top}")
