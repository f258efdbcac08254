#!/bin/sh
# quick A/B on the GPU: bench numbers for tile variants (MGRL_TILE=64|128)
for tile in 128 64; do
  echo "== MGRL_TILE=$tile"
  MGRL_TILE=$tile python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-ppo 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('  many=%.4e frac=%.4f | per-step-launch=%.3e us/launch=%.2f err=%s' % (j['value'], j['roofline']['frac'], j['per_step_launch']['value'], j['per_step_launch']['us_per_launch'], j['env_error_flags']))
"
done
