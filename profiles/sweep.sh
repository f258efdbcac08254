#!/bin/sh
# launch-shape sweep of the step kernel (environments per CTA x reset scheduling)
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits 2>&1 | head -3
for envs in 65536 1048576; do
for tile in 64 128 256; do for spread in 0 1; do
  echo "envs=$envs tile=$tile spread=$spread"
  MGRL_TILE=$tile MGRL_SPREAD=$spread python bench.py --envs $envs --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('  value=%.3e us/launch=%.2f frac=%.3f many=%.3e many_frac=%.3f clocks=%s' % (j['value'], j['roofline']['us_per_launch'], j['roofline']['frac'], j['step_many']['value'], j['step_many']['frac'], j['clocks']))
"
done; done; done
