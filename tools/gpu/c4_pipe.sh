#!/bin/bash
# c4: parity of the pipelined sorted kernel (sort_pipe) + bench variants given as HB_OPTIONS JSON strings (under gpurun, repo root)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "degenerate and pipe" 2>&1 | tail -5 || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sorted or tuning or fresh or half or 320" 2>&1 | tail -5
for o in "$@"; do
  echo "$o"
  HB_OPTIONS="$o" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   %.2f Gev/s  frac %.4f  %s' % (d['value']/1e9, d['roofline']['frac'], d['config']['kernel']))
    else: print(l[:300])
"
done
