#!/bin/bash
# Second (last) evidence run of round 2's third session: the defaults after the first run (split_table off, 128 lanes per
# group in the summing phase) against the paired summing loop (sort_pair) and the next-tile L2 prefetch (sort_prefetch).
mkdir -p gpurun_out
L=gpurun_out/r02_last2.log
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
         '{"sort_pair": true}' \
         '{"sort_prefetch": true}' \
         '{"sort_pair": true, "sort_prefetch": true}' \
         '{"sort_pair": true, "sort_prefetch": true, "sort_threads": 768, "sort_ept": 4}' \
         '{"sort_max_lanes": 64}' ; do
  say "$o"
  HB_OPTIONS="$o" timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu --no-extras --no-e2e 2>&1 | bench_line | tee -a $L
done
say "== parity of the sorted kernel with sort_pair + sort_prefetch (HB_OPTIONS)"
HB_OPTIONS='{"sort_pair": true, "sort_prefetch": true}' timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_gpu.py -m gpu -x -q -k "sorted or fresh or c4 or 320 or half" 2>&1 | tail -3 | tee -a $L
say "== ncu captures at 2e7 events: the defaults, then sort_pair + sort_prefetch (each after the same command exited 0 without ncu)"
python bench.py --events 2e7 --steps 3 --no-e2e --no-cpu --no-extras > gpurun_out/plain2.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:hb_fill_kernel -s 13 -c 2 -o gpurun_out/prof_c4_r02c python bench.py --events 2e7 --steps 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu3.log 2>&1
tail -1 gpurun_out/ncu3.log | tee -a $L
HB_OPTIONS='{"sort_pair": true, "sort_prefetch": true}' python bench.py --events 2e7 --steps 3 --no-e2e --no-cpu --no-extras > gpurun_out/plain3.log 2>&1 &&
HB_OPTIONS='{"sort_pair": true, "sort_prefetch": true}' timeout 200 ncu --set full --clock-control none --import-source on -k regex:hb_fill_kernel -s 13 -c 2 -o gpurun_out/prof_c4_r02d python bench.py --events 2e7 --steps 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu4.log 2>&1
tail -1 gpurun_out/ncu4.log | tee -a $L
say "== launch list of the default bench command"
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/plain.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c4_launches_b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu1.log 2>&1
say "== GPU tests, all files, the defaults"
timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $L
say "== done"
