"""Aggregates an ncu gpu__time_duration launch list (csv) by kernel name."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = {}
for r in rows:
    k = r[4].split("(")[0][-44:]
    agg.setdefault(k, [0.0, 0])
    agg[k][0] += float(r[-1]); agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
    print("%-46s %10.3f ms  x%-3d %5.1f%%" % (k, v[0] / 1e6, v[1], 100 * v[0] / tot))
