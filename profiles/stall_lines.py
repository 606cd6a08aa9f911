"""Aggregates `ncu --page source --csv --print-source cuda,sass` output: top source lines by warp-stall samples."""
import csv, sys
def I(x):
    try: return int(x)
    except Exception: return 0
rows=list(csv.reader(open(sys.argv[1])))
secs=[]; cur=None
for r in rows:
    if r and r[0]=='File Path':
        cur={'rows':[]}; secs.append(cur)
    elif cur is not None: cur['rows'].append(r)
for si,s in enumerate(secs):
    rr=s['rows']; name=rr[0][1] if rr and rr[0] else '?'; hdr=rr[1]; data=rr[2:]
    iS=hdr.index('# Samples')
    stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    lines=[r for r in data if r and r[0]!='' and len(r)>iS]
    tot=sum(I(r[iS]) for r in lines) or 1
    print('=== kernel',si,name,'total samples',tot)
    allst={}
    for r in lines:
        for i in stall_cols:
            if i<len(r): allst[hdr[i][6:]]=allst.get(hdr[i][6:],0)+I(r[i])
    print('   stall mix:', {k:round(100*v/tot,1) for k,v in sorted(allst.items(), key=lambda x:-x[1])[:7]})
    for r in sorted(lines,key=lambda r:-I(r[iS]))[:int(sys.argv[2]) if len(sys.argv)>2 else 12]:
        st={hdr[i][6:]:I(r[i]) for i in stall_cols if i<len(r) and I(r[i])>0}
        st=dict(sorted(st.items(), key=lambda x:-x[1])[:3])
        print(f"{100*I(r[iS])/tot:5.1f}%  L{r[0]:>4} {r[1].strip()[:100]}  {st}")
