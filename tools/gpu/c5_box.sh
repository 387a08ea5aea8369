#!/bin/bash
# c5 box window: parity + bench variants (run under gpurun from the repo root)
timeout 900 python -m pytest tests -m gpu -x -q -k "c5 or tuning or strategies or deterministic" 2>&1 | tail -4
for o in '{}' '{"box":false}' '{"win_pair":false}'; do
  echo "$o"
  HB_OPTIONS="$o" timeout 200 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu --no-extras --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value']/1e9, d['roofline']['frac'], d['config']['kernel'])
    else: print(l[:300])
"
done
