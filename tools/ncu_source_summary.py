import csv,re,sys
rows=list(csv.reader(open(sys.argv[1])))
cur=None; lines={}
for r in rows:
    if len(r)>=2 and r[0]=="File Path": cur=r[1]; continue
    if not r or not r[0].isdigit(): continue
    if r[2]=="-":
        try: lines[(cur.split('/')[-1], int(r[0]))]=(int(r[7]), int(r[4]), r[1].strip())
        except: pass
tot=sum(v[0] for v in lines.values()); ts=sum(v[1] for v in lines.values())
print("total inst", tot, "samples", ts)
src=open('/root/repo/tikv_b200/csrc/b2_device.h').read().split('\n')
funcs=[]
for i,l in enumerate(src,1):
    m=re.match(r'^(?:template <class V>\s*)?B2_HD\s+[\w:<>\* ]+?\s+(\w+)\(', l)
    if m: funcs.append((i,m.group(1)))
def func_of(line):
    name="?"
    for i,n in funcs:
        if i<=line: name=n
        else: break
    return name
from collections import Counter
ci=Counter(); cs=Counter()
for (f,l),(inst,smp,txt) in lines.items():
    k=('dev:'+func_of(l)) if f=='b2_device.h' else f+':'+str(l//20*20)
    ci[k]+=inst; cs[k]+=smp
for k,v in ci.most_common(22): print(f"{k:32s} inst {v:11d} {100*v/tot:5.1f}%  samples {cs[k]:6d} {100*cs[k]/ts:5.1f}%")
print("top lines by samples:")
for (f,l),(inst,smp,txt) in sorted(lines.items(), key=lambda kv:-kv[1][1])[:18]: print(f"  {f}:{l} inst {inst} smp {smp} | {txt[:110]}")
