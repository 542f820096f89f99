#!/usr/bin/env python
"""Times the stepping kernel alone (CUDA events on the handle's stream) for a few workloads.
Used for kernel-variant experiments:  MRSB_LIB_PATH=<variant.so> python tools/time_step.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import workload, x500_world  # noqa: E402
from mrs_multirotor_simulator_b200 import (ACTUATOR_CMD, POSITION_CMD, VELOCITY_HDG_CMD, VELOCITY_HDG_RATE_CMD, UavBatch,  # noqa: E402
                                           airframe)


def time_case(name, n, mode, k, reps=30, types=None, tou=None):
    spawn, cmd = workload(0, n)
    spawn[:, 2] = 10.0
    b = UavBatch(types or [x500_world()], type_of_uav=tou, spawn_xyz=spawn, n=n)
    if mode == ACTUATOR_CMD:
        cmd = np.full((n, 8), 0.55)
    elif mode == POSITION_CMD:
        cmd = np.concatenate([spawn[:, :2] + cmd[:, :2], 12.0 + cmd[:, 2:3], cmd[:, 3:4]], axis=1)
    b.set_input(mode, cmd)
    st = torch.cuda.ExternalStream(b.stream)
    for _ in range(5):
        b.make_step(0.01, k)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, e in ev:
        a.record(st)
        b.make_step(0.01, k)
        e.record(st)
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(e) for a, e in ev])
    return {"case": name, "n": n, "k": k, "ms_median": float(np.median(ms)), "ms_min": float(ms.min()),
            "uav_steps_per_s": n * k / (float(np.median(ms)) * 1e-3)}


if __name__ == "__main__":
    M = 1 << 20
    mixed = [airframe(f, ground_enabled=True) for f in ("x500", "f550", "naki")]
    cases = [("vel_hdg_rate 1M K=1", M, VELOCITY_HDG_RATE_CMD, 1, None, None), ("position 1M K=1", M, POSITION_CMD, 1, None, None),
             ("actuator 1M K=1", M, ACTUATOR_CMD, 1, None, None), ("vel_hdg 64k K=10", 65536, VELOCITY_HDG_CMD, 10, None, None),
             ("vel_hdg 1M K=10", M, VELOCITY_HDG_CMD, 10, None, None),
             ("actuator mixed 4/6/8 1M K=1", M, ACTUATOR_CMD, 1, mixed, (np.arange(M) % 3).astype(np.int32))]
    out = [time_case(c[0], c[1], c[2], c[3], types=c[4], tou=c[5]) for c in cases]
    print(json.dumps({"lib": os.environ.get("MRSB_LIB_PATH", "default"), "results": out}))
