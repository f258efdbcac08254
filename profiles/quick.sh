#!/bin/sh
# quick A/B on the GPU: tests (optional) + bench numbers for tile variants
for tile in 128 64; do
  echo "== MGRL_TILE=$tile"
  MGRL_TILE=$tile python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('  per-step-launch value=%.3e us/launch=%.2f | many=%.3e frac=%.3f err=%s' % (j['value'], j['roofline']['us_per_launch'], j['step_many']['value'], j['step_many']['frac'], j['env_error_flags']))
"
done
