#!/usr/bin/env python
"""Times the collision pass alone on the bench workload (positions after a short flight)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import workload, x500_world  # noqa: E402
from mrs_multirotor_simulator_b200 import VELOCITY_HDG_RATE_CMD, UavBatch  # noqa: E402

n = 1 << 20
spawn, cmd = workload(0, n)
b = UavBatch([x500_world()], spawn_xyz=spawn, n=n)
b.set_input(VELOCITY_HDG_RATE_CMD, cmd)
b.set_collisions(True, False, 100.0)
b.run(0.01, 300, with_collisions=True)
st = torch.cuda.ExternalStream(b.stream)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
torch.cuda.synchronize()
for a, e in ev:
    a.record(st)
    b.handle_collisions()
    e.record(st)
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(e) for a, e in ev])
print(json.dumps({"lib": os.environ.get("MRSB_LIB_PATH", "default"), "collision_pass_us": float(np.median(ms)) * 1000, "pairs": b.counters()["pairs"]}))
