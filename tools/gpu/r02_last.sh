#!/bin/bash
# Last evidence run of round 2 (under gpurun, one GPU, ~11 minutes of commands): the big-book kernel after the third session's
# changes (full-tile copy of the event code, group_fast variant, flag bits for filtered weights, split index from the bin index,
# odd chunks).  Order = importance: a call that is cut short still leaves the earlier parts in gpurun_out/.
mkdir -p gpurun_out
L=gpurun_out/r02_last.log
: > $L
say() { echo "$@" | tee -a $L; }
bench_line() {
python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); k = d['config'].get('kernel') or {}
        print('   %.3f Gev/s  %.3f ms/step  frac %.4f  regs %s smem %s grid %s' % (d['value'] / 1e9, d['ms_per_step'], d['roofline']['frac'], k.get('regs'), k.get('smem_bytes'), k.get('grid')))
    elif l.strip():
        print('   | ' + l.rstrip()[:260])
"
}
say "== smoke"
timeout 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee -a $L
say "== c4 variants (bench.py --steps 4 --warmup 3, device-resident columns, 1e9 events)"
for o in '{}' \
         '{"sort_chunk_odd": false, "sort_full": false, "split_table": false, "sort_gate": false, "group_fast": false}' \
         '{"split_table": false}' \
         '{"sort_lanes": 128}' \
         '{"sort_gate": false}' \
         '{"sort_full": false}' \
         '{"sort_threads": 768, "sort_ept": 4}' ; do
  say "$o"
  HB_OPTIONS="$o" timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu --no-extras --no-e2e 2>&1 | bench_line | tee -a $L
done
say "== c1 at 1e7 events, blocks per SM"
for o in '{}' '{"max_blocks_per_sm": 1}'; do
  say "$o"
  HB_OPTIONS="$o" timeout 120 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu --no-extras --no-e2e 2>&1 | bench_line | tee -a $L
done
say "== ncu capture of the default c4 kernel at 2e7 events (after the same command exited 0 without ncu)"
python bench.py --events 2e7 --steps 3 --no-e2e --no-cpu --no-extras > gpurun_out/plain2.log 2>&1 &&
timeout 240 ncu --set full --clock-control none --import-source on -k regex:hb_fill_kernel -s 13 -c 2 -o gpurun_out/prof_c4_r02b python bench.py --events 2e7 --steps 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | tee -a $L
say "== GPU tests: parity files first"
timeout 330 python -m pytest tests/test_gpu_parity.py tests/test_api_gpu.py tests/test_fullsize_gpu.py tests/test_random_definitions_gpu.py tests/test_deterministic_gpu.py tests/test_dist_gpu.py -m gpu -x -q --durations=8 2>&1 | tail -16 | tee -a $L
say "== done"
