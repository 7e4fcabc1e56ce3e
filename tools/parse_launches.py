#!/usr/bin/env python
"""One step of an ncu launch list (--metrics gpu__time_duration.sum --csv): per-launch times, per-kernel totals and shares.
Usage: python tools/parse_launches.py launches.csv [launches_per_step [first_launch_of_the_step]]"""
import csv, collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
per = int(sys.argv[2]) if len(sys.argv) > 2 else 43
skip = int(sys.argv[3]) if len(sys.argv) > 3 else per
rows=[]
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum':
        rows.append((int(row['ID']), row['Kernel Name'], float(row['Metric Value'].replace(',','')), row['Grid Size'], row['Block Size']))
step = rows[skip:skip+per]
tot=sum(v for _,_,v,_,_ in step)
for i,(id_,k,v,g,b) in enumerate(step):
    print(f"{i:3d} {k[:44]:44s} {v/1000:9.1f} us  grid {g} block {b}")
print("step total us", tot/1000, "launches", len(step), "of", len(rows))
agg=collections.Counter()
for _,k,v,_,_ in step: agg[k.split('(')[0]]+=v
for k,v in agg.most_common(): print(f"{k:30s} {v/1000:9.1f} us {100*v/tot:5.1f}%")
