import csv, collections, sys
path=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 60
rows=list(csv.reader(open(path)))
agg=collections.defaultdict(lambda:[0,0,0])
cur=None; N=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': N=len(r); continue
    if r[0]=='' or N is None: continue
    try: ln=int(r[0])
    except: continue
    extra=len(r)-N
    src=','.join(r[1:2+extra]); rest=r[2+extra:]
    # rest[0]=Address,rest[1]=Source(sass), rest[2]=all samples, rest[3]=not issued, rest[4]=#samples, rest[5]=inst, rest[6]=thread inst
    try:
        a=agg[(cur.split('/')[-1],ln,src.strip()[:100])]
        a[0]+=int(rest[5]); a[1]+=int(rest[6]); a[2]+=int(rest[4])
    except Exception as e: pass
tot=sum(a[0] for a in agg.values()); ts=sum(a[2] for a in agg.values())
print('total inst',tot,'samples',ts)
byfile=collections.Counter()
for k,a in agg.items(): byfile[k[0]]+=a[0]
print({k:round(100*v/tot,1) for k,v in byfile.items()})
for k,a in sorted(agg.items(), key=lambda x:-x[1][0])[:topn]:
    print(f"{100*a[0]/tot:5.2f}% inst {100*a[2]/ts:5.2f}% smp  lanes {a[1]/max(a[0],1):4.1f}  {k[0]}:{k[1]}  {k[2]}")
