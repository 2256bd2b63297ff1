import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
warp_rows=float(sys.argv[2]); lo=int(sys.argv[4]); hi=int(sys.argv[5]); fname=sys.argv[3]
cur=None
for r in rows:
    if len(r)>=2 and r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if not r or not r[0].isdigit(): continue
    if r[2]=="-" and cur==fname and lo<=int(r[0])<=hi:
        try:
            inst=int(r[7]); smp=int(r[4])
        except: continue
        if inst: print(f"{r[0]:>5} {inst/warp_rows:7.1f} smp {smp:5d} | {r[1].strip()[:130]}")
