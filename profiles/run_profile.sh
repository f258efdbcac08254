#!/bin/sh
# Run under gpurun from the repo root:  gpurun --timeout 1500 -- 'sh profiles/run_profile.sh r08'
# 1) plain bench (the number), 2) ncu launch list of the bench command's kernel path (prof_target.py issues the same
#    launches deterministically), 3) one `--set full` capture per kernel.  Outputs land in gpurun_out/.
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
BENCH="python bench.py --steps 2 --warmup 3 --no-ppo --no-e2e --no-cpu-baseline"
$BENCH > gpurun_out/plain_bench_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_bench_$TAG.csv $BENCH > gpurun_out/ncu_launches_bench_$TAG.log 2>&1
echo "ncu launches bench rc=$?"
for MODE in ${MODES:-many step ppo gae}; do
  python profiles/prof_target.py $MODE > gpurun_out/plain_${MODE}_$TAG.log 2>&1 || { echo "plain $MODE failed"; continue; }
  case $MODE in
    many) K=rollout_kernel; SKIP=3; CNT=1;;
    gae)  K=gae_kernel; SKIP=3; CNT=1;;
    step) K=step_kernel; SKIP=400; CNT=1;;
    ppo)  K=policy_forward; SKIP=40; CNT=1;;
  esac
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT \
      -o gpurun_out/prof_${MODE}_$TAG -f python profiles/prof_target.py $MODE > gpurun_out/ncu_full_${MODE}_$TAG.log 2>&1
  echo "ncu full $MODE rc=$?"
done
ls gpurun_out | grep $TAG
