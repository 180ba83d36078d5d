"""Summarise `ncu --page source --csv` of one launch: executed warp instructions and stall samples by code region.
usage: ncu -i X.ncu-rep --page source --csv --launch-skip N --launch-count 1 | python scripts/ncu_src_summary.py [seg]"""
import csv, sys
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 250
rows = list(csv.reader(sys.stdin))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
ie = ix['Instructions Executed']; isamp = ix['# Samples']
tot = sum(int(d[ie]) for d in data); ts = sum(int(d[isamp]) for d in data)
print('kernel', rows[0][1], 'warp instructions', tot, 'samples', ts, 'SASS lines', len(data))
for s in range(0, len(data), seg):
    e = sum(int(d[ie]) for d in data[s:s + seg]); sm = sum(int(d[isamp]) for d in data[s:s + seg])
    if e / tot > 0.005 or sm / ts > 0.005:
        print(f"{s:6d} exec {e / tot * 100:5.1f}%  samples {sm / ts * 100:5.1f}%  {data[s][1].strip()[:60]}")
