#!/usr/bin/env python
"""Under torchrun: per-tick device times of the stepping launch and of the collision pass on every rank of a sharded run of
the bench swarm (CUDA events on the handle's stream, nothing synchronises inside the loop).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/time_multi.py

N_UAVS (default 1 Mi), TICKS (default 400), USE_RUN=1 (drive the ticks through mrsb_run: one graph per tick)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import workload, x500_world  # noqa: E402
from mrs_multirotor_simulator_b200 import ACTUATOR_CMD, VELOCITY_HDG_RATE_CMD, UavBatch  # noqa: E402
from mrs_multirotor_simulator_b200.sharding import connect, shard_range  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(os.environ.get("N_UAVS", 1 << 20))
ticks = int(os.environ.get("TICKS", 400))
begin, count = shard_range(n, world, rank)
spawn, cmd = workload(begin, count)
b = UavBatch([x500_world()], spawn_xyz=spawn, n=count, device=local, n_global=n, shard_begin=begin)
if world > 1:
    connect(b, dist)
b.set_input(ACTUATOR_CMD, np.zeros((count, 8)))
b.make_step(0.01)
b.make_step(0.01)
b.set_collisions(True, False, 100.0)
b.set_input(VELOCITY_HDG_RATE_CMD, cmd)
b.run(0.01, int(os.environ.get("FAST_FORWARD", 300)))
b.sync()
st = torch.cuda.ExternalStream(b.stream, device=torch.device("cuda", local))
if world > 1:
    dist.barrier()
out = {"rank": rank, "world": world, "n_local": count, "exchange_mode": b.exchange_mode()}
if os.environ.get("USE_RUN"):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0 = b.collision_info()
    a.record(st)
    b.run(0.01, ticks)
    e.record(st)
    b.sync()
    i1 = b.collision_info()
    out.update({"tick_us": a.elapsed_time(e) * 1000 / ticks, "rebuild_fraction": (i1["rebuilds"] - i0["rebuilds"]) / ticks})
else:
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(ticks)]
    i0 = b.collision_info()
    for t in range(ticks):
        ev[t][0].record(st)
        b.make_step(0.01)
        ev[t][1].record(st)
        b.handle_collisions()
        ev[t][2].record(st)
    b.sync()
    i1 = b.collision_info()
    step = np.array([e[0].elapsed_time(e[1]) for e in ev]) * 1000
    coll = np.array([e[1].elapsed_time(e[2]) for e in ev]) * 1000
    total = ev[0][0].elapsed_time(ev[-1][2]) * 1000 / ticks
    k = max(1, int(round((i1["rebuilds"] - i0["rebuilds"]))))
    order = np.sort(coll)
    out.update({"tick_us": total, "step_us_median": float(np.median(step)), "pass_us_median": float(np.median(coll)), "pass_us_p90": float(np.quantile(coll, 0.9)),
                "pass_us_mean": float(coll.mean()), "pass_us_mean_of_rebuild_ticks": float(order[-k:].mean()), "pass_us_mean_of_list_ticks": float(order[:-k].mean()),
                "rebuild_fraction": (i1["rebuilds"] - i0["rebuilds"]) / ticks, "pairs_last": b.counters()["pairs"]})
if os.environ.get("MRSB_TIMELINE"):
    tl = b.timeline(ticks).astype(np.int64)
    if len(tl) > 10:
        lst = tl[tl[:, 5] == 0]
        d = lambda a: float(np.median(a)) / 1000.0
        nxt = tl[1:, 0] - tl[:-1, 4]
        keep = tl[:-1, 5] == 0
        out["timeline_us_list_ticks"] = {"signal": d(lst[:, 1] - lst[:, 0]), "wait_for_peers": d(lst[:, 2] - lst[:, 1]), "decide_to_refresh": d(lst[:, 3] - lst[:, 2]),
                                         "refresh_to_check": d(lst[:, 4] - lst[:, 3]), "check_start_to_next_pass_start(check+step+gaps)": d(nxt[keep]),
                                         "pass_start_period": d(tl[1:, 0] - tl[:-1, 0])}
        reb = tl[tl[:, 5] == 1]
        if len(reb):
            out["timeline_us_rebuild_ticks"] = {"decide_to_list_build": d(reb[:, 7] - reb[:, 2]), "list_build_to_check": d(reb[:, 4] - reb[:, 7]), "n": int(len(reb))}
if world > 1:
    res = [None] * world
    dist.all_gather_object(res, out)
else:
    res = [out]
if rank == 0:
    for r in res:
        print(json.dumps(r), flush=True)
if world > 1:
    dist.barrier()
    b.close()
    dist.destroy_process_group()
