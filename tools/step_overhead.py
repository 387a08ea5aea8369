#!/usr/bin/env python
"""Where the host time of one benchmark step goes: Book.fill of c4 on a small shard + dist.Merger.start() (world = 1, NCCL).
    python tools/step_overhead.py [events]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
import torch, torch.distributed as dist
import bench
from histbook_b200 import dist as hbdist, engine

engine.init(0)
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**6
cols = bench.device_columns(torch, "c4", n, 1, torch.device("cuda", 0))
target = bench.Target("c4")
merger = hbdist.Merger(target.obj)
for _ in range(5):
    target.fill(cols); merger.start()
merger.wait(); torch.cuda.synchronize()
reps = 50
t0 = time.perf_counter()
for _ in range(reps):
    target.fill(cols)
torch.cuda.synchronize(); t1 = time.perf_counter()
for _ in range(reps):
    merger.start()
merger.wait(); torch.cuda.synchronize(); t2 = time.perf_counter()
for _ in range(reps):
    target.fill(cols); merger.start()
merger.wait(); torch.cuda.synchronize(); t3 = time.perf_counter()
print("events %d: Book.fill %.3f ms, Merger.start %.3f ms, both %.3f ms per step" % (n, 1e3 * (t1 - t0) / reps, 1e3 * (t2 - t1) / reps, 1e3 * (t3 - t2) / reps))
pr = cProfile.Profile(); pr.enable()
for _ in range(reps):
    target.fill(cols); merger.start()
merger.wait(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
dist.destroy_process_group()
