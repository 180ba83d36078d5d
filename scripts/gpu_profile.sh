#!/bin/bash
# Run on the B200 box (gpurun): GPU parity tests, one bench line, the ncu launch list of the same
# command and one full ncu capture of the dominant kernel.  Outputs land in gpurun_out/.
# Usage: scripts/gpu_profile.sh <tag> [kernel-regex]
TAG=${1:-r1}
KRE=${2:-k_tc_linear}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
# launch list: only this library's kernels (all named k_*); 61 launches per F2 step (1 opt_init, 20 chain_pass and 19 chain_grad with the keep-best bookkeeping fused in, 19 adam, 1 keepbest_b, 1 finalize with the last snapshot folded in): skip plan creation + the 3 warm-up
# steps (-s counts launches that match -k), list the 2 timed steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s ${LSKIP:-193} -c ${LCOUNT:-122} --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --only-f2 --no-profile \
    > gpurun_out/${TAG}_launches.log 2>&1
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${SKIP:-20} -c ${COUNT:-2} \
    -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --only-f2 --no-profile \
    > gpurun_out/${TAG}_prof.log 2>&1
ls -la gpurun_out/
