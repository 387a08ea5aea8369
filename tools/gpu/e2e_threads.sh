#!/bin/bash
# e2e (pageable host columns) against the number of copy threads of the bounce ring
for t in 4 8 12 16; do
  echo "HB_COPY_THREADS=$t"
  HB_COPY_THREADS=$t timeout 300 python bench.py --steps 3 --warmup 3 --events 2e8 --e2e-events 2e8 --no-cpu --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('   pageable %.3f Gev/s (%.1f GB/s)   pinned %.3f Gev/s (%.1f GB/s)' % (e['value']/1e9, e['h2d_gbs_per_gpu'], e['pinned']['value']/1e9, e['pinned']['h2d_gbs_per_gpu']))
"
done
nproc; free -g | head -2
