#!/usr/bin/env python
"""Per-SASS-instruction stall samples of an ncu capture (--import-source on):
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv ; python profiles/ncu_sass.py x.csv [top]
Prints the stall-reason totals of the kernel and the instructions with the most samples (with their dominant reasons)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
full = max((r for r in rows if r and r[0] == "Line No"), key=len)
reasons = [i for i, h in enumerate(full) if h.startswith("stall_") and "Not Issued" not in h]
tot = {}
ins = []
for r in rows:
    if len(r) > 4 and r[2].startswith("0x"):
        try:
            s = int(r[4])
        except ValueError:
            continue
        rs = {}
        for i in reasons:
            if i < len(r) and r[i] not in ("", "-"):
                try:
                    v = int(r[i])
                except ValueError:
                    continue
                if v:
                    rs[full[i]] = v
                    tot[full[i]] = tot.get(full[i], 0) + v
        ex = r[7] if len(r) > 7 else ""
        ins.append((s, r[2], r[3].strip(), ex, rs))
T = sum(tot.values())
print("stall samples by reason (all samples): total", T)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("  %-26s %6.2f%%" % (k, 100.0 * v / T))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("top instructions:")
for s, a, op, ex, rs in sorted(ins, key=lambda x: -x[0])[:top]:
    why = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:3])
    print("  %7d  %s  %-50s exec=%s  [%s]" % (s, a[-6:], op[:50], ex, why))
