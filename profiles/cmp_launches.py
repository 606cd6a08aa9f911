"""Per-layer comparison of two ncu launch lists (gpu__time_duration) of one bench step."""
import csv, re, sys
def load(fn):
    rows=list(csv.reader(open(fn)))
    hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    hdr=rows[hi]; data=rows[hi+1:]
    ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
    out=[]
    for r in data:
        if len(r)<=vi: continue
        v=float(r[vi].replace(',',''))
        if r[ui]=='ns': v/=1e3
        elif r[ui]=='ms': v*=1e3
        out.append((re.sub(r'\(.*','',r[ki]),v))
    return out
a=[x for x in load(sys.argv[1]) if 'umma' in x[0]]
b=[x for x in load(sys.argv[2]) if 'umma' in x[0]]
names=['up0']+[f's0.{i}' for i in range(18)]+['up1']+[f's1.{i}' for i in range(18)]+['up2']+[f's2.{i}' for i in range(18)]+['up3']+[f's3.{i}' for i in range(18)]
grp={}
for n,(x,y) in zip(names,zip(a,b)):
    g=n.split('.')[0]
    if len(sys.argv)>3: print(f"{n:6s} A {x[1]:8.1f}  B {y[1]:8.1f}  ratio {y[1]/x[1]:.2f}")
    ga=grp.setdefault(g,[0,0,0]); ga[0]+=x[1]; ga[1]+=y[1]; ga[2]+=min(x[1],y[1])
for g,(x,y,m) in grp.items(): print(f"{g:4s} A {x/1e3:7.2f} ms  B {y/1e3:7.2f} ms  min {m/1e3:7.2f}")
print('total A', sum(x[1] for x in a)/1e3, 'B', sum(x[1] for x in b)/1e3, 'min', sum(min(x[1],y[1]) for x,y in zip(a,b))/1e3)
