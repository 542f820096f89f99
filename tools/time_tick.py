#!/usr/bin/env python
"""Times whole ticks (stepping launch + collision pass) on the bench workload and reports how the
collision pass was organised (table rebuilds vs list checks).  MRSB_NO_NEIGHBOUR_LISTS=1 /
MRSB_COLLISION_CELL=<m> select the variants."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import workload, x500_world  # noqa: E402
from mrs_multirotor_simulator_b200 import VELOCITY_HDG_RATE_CMD, UavBatch  # noqa: E402

n = int(os.environ.get("N_UAVS", 1 << 20))
ticks = int(os.environ.get("TICKS", 1000))
spawn, cmd = workload(0, n)
b = UavBatch([x500_world()], spawn_xyz=spawn, n=n)
b.set_input(VELOCITY_HDG_RATE_CMD, cmd)
b.set_collisions(True, False, 100.0)
b.run(0.01, 20, with_collisions=True)
st = torch.cuda.ExternalStream(b.stream)
out = {"lists": os.environ.get("MRSB_NO_NEIGHBOUR_LISTS") is None, "cell_env": os.environ.get("MRSB_COLLISION_CELL")}
i0 = b.collision_info()
for label, count in (("first", ticks), ("second", ticks)):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(st)
    b.run(0.01, count, with_collisions=True)
    e.record(st)
    torch.cuda.synchronize()
    i1 = b.collision_info()
    out[label] = {"tick_us": a.elapsed_time(e) * 1000 / count, "rebuild_fraction": (i1["rebuilds"] - i0["rebuilds"]) / max(1, i1["passes"] - i0["passes"]),
                  "crowded_uavs": i1["crowded_uavs"], "pairs_last": b.counters()["pairs"]}
    i0 = i1
out["info"] = b.collision_info()
# step kernel alone
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
for a, e in ev:
    a.record(st)
    b.make_step(0.01)
    e.record(st)
torch.cuda.synchronize()
out["step_us"] = float(np.median([a.elapsed_time(e) for a, e in ev])) * 1000
out["collision_us_avg"] = out["second"]["tick_us"] - out["step_us"]
print(json.dumps({k: out[k] for k in ("step_us", "collision_us_avg", "first", "second", "lists", "cell_env")}))
if os.environ.get("VERBOSE"):
    print(json.dumps(out))
