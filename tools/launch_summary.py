"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel time of the last kick."""
import csv, sys
path = sys.argv[1]; kicks = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols = rows[hdr]; data = rows[hdr + 1:]
ki = cols.index('Kernel Name'); vi = cols.index('Metric Value'); ui = cols.index('Metric Unit')
names = []
for r in data:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    v = v / 1000.0 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1000.0)
    names.append((r[ki], v))
per = len(names) // kicks
last = names[-per:]
tot = sum(v for _, v in last)
for n, v in last:
    print(f"{v:9.2f} us {100*v/tot:5.1f}%  {n[:100]}")
print(f"total {tot:.1f} us over {per} launches")
