#!/usr/bin/env python
"""Stepping kernel alone (VelocityHdgRate, K = 1) at shard sizes, CUDA events, L2 flushed between launches when SIZES_FLUSH=1.
MRSB_LIB_PATH=<variant.so> selects a kernel-variant build."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import workload, x500_world  # noqa: E402
from mrs_multirotor_simulator_b200 import VELOCITY_HDG_RATE_CMD, UavBatch  # noqa: E402

flush = bool(os.environ.get("SIZES_FLUSH"))
buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None
out = {"lib": os.environ.get("MRSB_LIB_PATH", "default"), "flush": flush, "us": {}}
for n in (131072, 262144, 524288, 1048576):
    spawn, cmd = workload(0, n)
    spawn[:, 2] = 10.0
    b = UavBatch([x500_world()], spawn_xyz=spawn, n=n)
    b.set_input(VELOCITY_HDG_RATE_CMD, cmd)
    st = torch.cuda.ExternalStream(b.stream)
    for _ in range(20):
        b.make_step(0.01)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
    torch.cuda.synchronize()
    for a, e in ev:
        if flush:
            with torch.cuda.stream(st):
                buf.zero_()
        a.record(st)
        b.make_step(0.01)
        e.record(st)
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(e) for a, e in ev])
    out["us"][n] = {"median": round(float(np.median(ms)) * 1000, 2), "min": round(float(ms.min()) * 1000, 2), "grid": b.step_info()["grid"]}
    b.close()
print(json.dumps(out))
