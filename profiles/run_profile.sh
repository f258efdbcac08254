#!/bin/sh
# Run under gpurun from the repo root:  gpurun --timeout 1500 -- 'sh profiles/run_profile.sh r01'
# 1) plain bench (the number), 2) ncu launch list of the same short command, 3) one full capture
#    of the step kernel.  Outputs land in gpurun_out/ (copy summaries into profiles/).
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
SHORT="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$SHORT > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 384 -c 140 --csv \
    --log-file gpurun_out/launches_$TAG.csv $SHORT > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
$SHORT > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 400 -c 2 \
    -o gpurun_out/prof_$TAG -f $SHORT > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
