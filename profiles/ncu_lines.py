#!/usr/bin/env python
"""Per-source-line stall samples of an ncu capture taken with --import-source on (-lineinfo build):
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv ; python profiles/ncu_lines.py x.csv [top]"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
cur=None; agg={}
hdr=None
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>=2 and r[0]=='Line No': hdr=r; continue
    if len(r)>4 and r[0] not in ('','Function Name') and r[0].isdigit():
        try: s=int(r[4])
        except: continue
        ex=0
        try: ex=int(r[hdr.index('Instructions Executed')])
        except: pass
        agg[(cur,int(r[0]))]=(agg.get((cur,int(r[0])),(0,0,''))[0]+s, agg.get((cur,int(r[0])),(0,0,''))[1]+ex, r[1])
tot=sum(v[0] for v in agg.values())
print("total samples",tot)
top=sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[2]) if len(sys.argv)>2 else 40]
for (f,l),(s,ex,src) in top:
    print(f"{100*s/tot:5.1f}% {s:7d} inst={ex:9d} {f}:{l}  {src.strip()[:110]}")
