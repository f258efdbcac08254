#!/bin/sh
# e2e (mgrl_vec_step_frames_host) with the AVX-512 VBMI / SSSE3 expansion, for several host-thread counts; run under gpurun
grep -m1 "model name" /proc/cpuinfo; grep -o "avx512vbmi" /proc/cpuinfo | sort | uniq -c
B="python bench.py --steps 3 --warmup 3 --no-ppo --no-cpu-baseline --no-configs"
for th in ${THREADS:-16 8 4 2}; do
  for isa in avx512 ssse3; do
    if [ $isa = ssse3 ]; then export MGRL_WIRE_NO_AVX512=1; else unset MGRL_WIRE_NO_AVX512; fi
    MGRL_WIRE_DEBUG=1 MGRL_HOST_THREADS=$th $B 2>/tmp/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads=$th $isa e2e %.1f M env-steps/s, stacked %.1f M' % (d['e2e']['value']/1e6, d['e2e']['stacked']['value']/1e6))"
    grep "mgrl_wire. frames " /tmp/err.txt | head -1
  done
done
