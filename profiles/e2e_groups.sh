#!/bin/sh
# e2e (mgrl_vec_step_frames_host): AVX-512 expansion in groups of 16 records (register realignment, vector scalars) against the
# staged one-record loop (MGRL_WIRE_NO_GROUPS=1); run under gpurun
B="python bench.py --steps 3 --warmup 3 --no-ppo --no-cpu-baseline --no-configs"
for g in 0 1 0 1; do
  if [ $g = 1 ]; then export MGRL_WIRE_NO_GROUPS=1; else unset MGRL_WIRE_NO_GROUPS; fi
  MGRL_WIRE_DEBUG=1 $B 2>/tmp/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no_groups=$g e2e %.1f M env-steps/s, stacked %.1f M' % (d['e2e']['value']/1e6, d['e2e']['stacked']['value']/1e6))"
  grep "mgrl_wire. frames " /tmp/err.txt | tail -1
done
