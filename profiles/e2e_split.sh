#!/bin/sh
# e2e (mgrl_vec_step_frames_host) against the share of the batch copied directly, for several host-thread counts
# (MGRL_HOST_THREADS emulates ranks that share the host's cores); run under gpurun
B="python bench.py --steps 3 --warmup 3 --no-ppo --no-cpu-baseline --no-configs"
for th in ${THREADS:-16 4 2}; do
  for d in ${SPLITS:-0 auto 0.25 0.5 0.75}; do
    MGRL_WIRE_DEBUG=1 MGRL_HOST_THREADS=$th MGRL_WIRE_DIRECT=$d $B 2>/tmp/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads=$th direct=$d e2e %.1f M env-steps/s' % (d['e2e']['value']/1e6))"
    grep "mgrl_wire. frames " /tmp/err.txt | head -1
  done
done
