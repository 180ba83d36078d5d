O=gpurun_out; TAG=r2m
python bench.py --only-f2 --steps 1 --warmup 1 > $O/${TAG}_choices.json 2>/dev/null
export CROWN_B200_CONV_CHOICES=$(python -c "import json; print(json.loads(open('$O/${TAG}_choices.json').read().strip().splitlines()[-1])['plan']['conv_choices'])")
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file $O/${TAG}_launches_all.csv python bench.py --steps 1 --warmup 1 --only-f2 --no-profile > $O/${TAG}_launches.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open('$O/${TAG}_launches_all.csv', errors='ignore')))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[hdr_i], rows[hdr_i + 1:]
kn = hdr.index('Kernel Name')
seen, start = 0, 0
for i in range(len(data) - 1, -1, -1):
    if 'k_conv_tc(cb::ConvTcArgs)' in data[i][kn]:
        seen += 1
        if seen == 332:
            start = i
            break
j = start
while start > 0 and not any(t in data[start - 1][kn] for t in ('k_finalize', 'k_conv_tc_pack_w', 'k_conv_relayout')) and start > j - 40:
    start -= 1
with open('$O/${TAG}_launches.csv', 'w', newline='') as f:
    w = csv.writer(f); w.writerow(hdr); w.writerows(data[start:])
print('launches kept:', len(data) - start)
PY
rm -f $O/${TAG}_launches_all.csv
