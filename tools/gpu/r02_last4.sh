#!/bin/bash
# Final defaults of the round (sort_prefetch on): smoke + the default c4 bench line without the CPU / e2e / extras legs.
mkdir -p gpurun_out
L=gpurun_out/r02_last4.log
: > $L
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $L
timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-e2e 2>&1 | tail -1 | tee gpurun_out/r02_bench_c4_final.json | cut -c1-400 | tee -a $L
