"""Per-phase timing of one sharded sem_sim_join step (torchrun, B2_SHARD_TIMING=1): where the non-kernel time goes."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lotus_b200 import _native as nv  # noqa: E402
from lotus_b200.distributed import ShardedIndex, shard_bounds  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
os.environ["B2_SHARD_TIMING"] = "1"
dist.init_process_group("nccl", device_id=dev)
n, nq, d, k = 1_000_000, 100_000, 768, 32
lo, hi = shard_bounds(n, world, rank)
x = bench.gen_rows_torch(torch, lo, hi, d, 0, dev, torch.bfloat16)
q = bench.gen_rows_torch(torch, 0, nq, d, 1, dev, torch.bfloat16)
sh = ShardedIndex(x, lo)
for _ in range(3):
    sh.search(q, k)
acc = {}
steps = 5
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sh.search(q, k)
    e1.record()
    torch.cuda.synchronize()
    ph = dict(sh.last_phase_ms)
    ph["step_total"] = e0.elapsed_time(e1)
    for kk, v in ph.items():
        acc[kk] = acc.get(kk, 0.0) + v / steps
wall = (time.perf_counter() - t0) / steps * 1e3
acc["local search_dev minus filter (finalize, flags, sync)"] = acc["step_total"] - acc["filter"] - acc["all_gather(+wait for slowest rank)"] - acc["merge"]
print(f"[rank {rank}] wall {wall:.2f} ms/step | " + " | ".join(f"{kk}: {v:.2f}" for kk, v in acc.items()), flush=True)
dist.destroy_process_group()
