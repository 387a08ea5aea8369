#!/bin/bash
# Third (short) evidence run: 768 threads x 4 events with the next-tile prefetch, without the paired loop -- bench + parity subset.
mkdir -p gpurun_out
L=gpurun_out/r02_last3.log
: > $L
say() { echo "$@" | tee -a $L; }
O='{"sort_prefetch": true, "sort_threads": 768, "sort_ept": 4}'
say "$O"
HB_OPTIONS="$O" timeout 100 python bench.py --steps 4 --warmup 3 --no-cpu --no-extras --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); k = d['config'].get('kernel') or {}
        print('   %.3f Gev/s  %.3f ms/step  frac %.4f  regs %s smem %s grid %s' % (d['value'] / 1e9, d['ms_per_step'], d['roofline']['frac'], k.get('regs'), k.get('smem_bytes'), k.get('grid')))
    elif l.strip():
        print('   | ' + l.rstrip()[:260])
" | tee -a $L
say "== parity subset with these options"
HB_OPTIONS="$O" timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_gpu.py -m gpu -x -q -k "sorted or fresh or c4 or 320 or half" 2>&1 | tail -3 | tee -a $L
say "== done"
