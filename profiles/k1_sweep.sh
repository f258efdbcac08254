#!/bin/sh
# K1 tile-shape sweep (run under gpurun): env-steps/s of mgrl_step_many for step/generator warp counts per tile
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-ppo --no-e2e --no-cpu-baseline --no-configs"
for cfg in ${SWEEP:-14:8 14:6 14:4 7:4 7:3 7:2 4:2 2:1}; do
  sw=${cfg%%:*}; gw=${cfg##*:}
  MGRL_SW=$sw MGRL_GW=$gw $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sw=$sw gw=$gw', round(d['value']/1e9,3), 'G env-steps/s', round(d['roofline']['frac'],4), 'err', d['env_error_flags'])"
done
MGRL_ROLLOUT=0 $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('old step_kernel', round(d['value']/1e9,3), 'G env-steps/s')"
