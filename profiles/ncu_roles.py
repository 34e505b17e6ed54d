#!/usr/bin/env python
"""Stall samples of the DP kernel by warp role (loader warp vs row warps), from a source-page csv of an ncu capture:
   python profiles/ncu_roles.py x.csv <first line of v2_loader_group> <first line of v2_query>
A warp's samples are proportional to the time it is resident; the share that is NOT 'barrier' is the time it is busy."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
l0, l1 = int(sys.argv[2]), int(sys.argv[3])
full = max((r for r in rows if r and r[0] == "Line No"), key=len)
bi = full.index("stall_barrier")
cur = None; lineno = None; seen = set(); agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0].isdigit(): lineno = int(r[0])
    if len(r) > 4 and r[2].startswith("0x"):
        if r[2] in seen: continue
        seen.add(r[2])
        try: s = int(r[4])
        except ValueError: continue
        b = int(r[bi]) if bi < len(r) and r[bi] not in ("", "-") else 0
        role = "loader warp" if (cur == "mesh.cu" and l0 <= lineno < l1) else "row warps + rest"
        a = agg.setdefault(role, [0, 0]); a[0] += s; a[1] += b
tot = sum(v[0] for v in agg.values())
for k, (s, b) in agg.items():
    print("%-18s samples %7d (%.1f%% of all)  barrier %7d  busy %.1f%%" % (k, s, 100.0 * s / tot, b, 100.0 * (s - b) / s))
