#!/bin/sh
# Run under gpurun from the repo root:  gpurun --timeout 1500 -- 'sh profiles/run_profile.sh r03'
# 1) plain bench (the number), 2) ncu launch list of the profiling target, 3) one full capture of the
#    step kernel in each mode.  Outputs land in gpurun_out/ (copy summaries into profiles/).
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
for MODE in many step; do
  python profiles/prof_target.py $MODE > gpurun_out/plain_${MODE}_$TAG.log 2>&1 || { echo "plain $MODE failed"; continue; }
  if [ $MODE = many ]; then SKIP=3; CNT=2; LS=0; LC=8; else SKIP=400; CNT=2; LS=384; LC=128; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_kernel -s $LS -c $LC --csv \
      --log-file gpurun_out/launches_${MODE}_$TAG.csv python profiles/prof_target.py $MODE > gpurun_out/ncu_launches_${MODE}_$TAG.log 2>&1
  echo "ncu launches $MODE rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $SKIP -c $CNT \
      -o gpurun_out/prof_${MODE}_$TAG -f python profiles/prof_target.py $MODE > gpurun_out/ncu_full_${MODE}_$TAG.log 2>&1
  echo "ncu full $MODE rc=$?"
done
ls -la gpurun_out
