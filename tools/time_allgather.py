#!/usr/bin/env python
"""Times NCCL all-gather of the packed positions (24 B x 1 Mi UAVs total) — the per-tick exchange."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 3 * (1 << 20)
buf = torch.zeros(n, dtype=torch.float64, device="cuda")
mine = buf[rank * n // world:(rank + 1) * n // world]
for _ in range(10):
    dist.all_gather_into_tensor(buf, mine)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
for a, b in ev:
    a.record()
    dist.all_gather_into_tensor(buf, mine)
    b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
if rank == 0:
    print(f"world {world}: all-gather {n * 8 / 1e6:.1f} MB total: median {ms[len(ms) // 2] * 1000:.1f} us, min {ms[0] * 1000:.1f} us")
dist.destroy_process_group()
