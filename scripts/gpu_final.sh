#!/bin/bash
# Round-end measurement on the B200 box (gpurun): GPU parity tests, the bench lines, the ncu launch list of the bench
# command and one `ncu --set full` capture of the dominant kernel class (k_conv_tc) of the headline workload.
# Usage: scripts/gpu_final.sh <tag>
TAG=${1:-r2final}
O=gpurun_out
mkdir -p $O
if [ -z "$PROF_ONLY" ]; then
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err
fi
# the profiled runs replay the convolution kernel choices of the un-profiled bench (see cb_plan_conv_choices)
python bench.py --only-f2 --steps 1 --warmup 1 > $O/${TAG}_choices.json 2>/dev/null
export CROWN_B200_CONV_CHOICES=$(python -c "import json,sys; print(json.loads(open('$O/${TAG}_choices.json').read().strip().splitlines()[-1])['plan']['conv_choices'])")
echo "conv choices: $CROWN_B200_CONV_CHOICES"
# launch list of the bench's timed loop (this library's kernels are all named k_*)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv \
    --log-file $O/${TAG}_launches_all.csv python bench.py --steps 1 --warmup 1 --only-f2 --no-profile > $O/${TAG}_launches.log 2>&1
N=$(grep -c 'k_conv_tc(cb::ConvTcArgs)' $O/${TAG}_launches_all.csv)
SKIP=$((N-332))
echo "k_conv_tc launches: $N, one step = 332, skipping $SKIP" | tee -a $O/${TAG}_launches.log
# keep the last step only in the committed list (plan creation times every conv layer, the warm-up step repeats the step)
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
# the step begins a few launches before its first convolution (opt_init, the first layers of the pass)
while start > 0 and not any(t in data[start - 1][kn] for t in ('k_finalize', 'k_conv_tc_pack_w', 'k_conv_relayout')) and start > i - 40:
    start -= 1
with open('$O/${TAG}_launches.csv', 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(hdr)
    w.writerows(data[start:])
print('launches kept:', len(data) - start)
PY
rm -f $O/${TAG}_launches_all.csv
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^k_conv_tc$" -s $SKIP -c 18 \
    -o $O/${TAG}_conv_tc -f python bench.py --steps 1 --warmup 1 --only-f2 --no-profile > $O/${TAG}_prof.log 2>&1
tail -2 $O/${TAG}_prof.log
ls -la $O/ | tail -15
