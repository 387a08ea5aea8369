#!/bin/bash
# round-2 evidence run (under gpurun, one GPU): tests, default bench, launch list, full capture of the top kernels
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_tests.log; cat gpurun_out/r02_tests.log; fi
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.err; cut -c1-200 gpurun_out/r02_bench_n1.json
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c4_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu1.log 2>&1
python bench.py --events 2e7 --steps 3 --no-e2e --no-cpu --no-extras > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hb_fill_kernel -s 13 -c 2 -o gpurun_out/prof_c4_r02 python bench.py --events 2e7 --steps 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log
