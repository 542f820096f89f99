#!/usr/bin/env python
"""Error statistics of the GPU stepping path against the CPU oracle after 10 s of flight, per input mode: the distribution over
UAVs of the largest component error (absolute, and relative to the excursion for the open-loop modes whose flight is chaotic).
Used to judge arithmetic variants of the kernel:  MRSB_LIB_PATH=<variant.so> python tools/err_stats.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import grid_spawn, make_pair, rand  # noqa: E402
from oracle import binding as O  # noqa: E402
from test_step_parity import _commands  # noqa: E402
from mrs_multirotor_simulator_b200 import airframe  # noqa: E402

n = int(os.environ.get("N", 4096))
out = {"lib": os.environ.get("MRSB_LIB_PATH", "default"), "n": n}
for mode, name in ((O.ACTUATOR_CMD, "actuator"), (O.CONTROL_GROUP_CMD, "control_group"), (O.ATTITUDE_RATE_CMD, "attitude_rate"),
                   (O.VELOCITY_HDG_RATE_CMD, "velocity_hdg_rate"), (O.POSITION_CMD, "position")):
    orc, gpu = make_pair([airframe("x500")], None, grid_spawn(n, z=10.0), rand(3, 0, n, -3, 3))
    cmd = _commands(mode, n)
    orc.set_input(mode, cmd)
    gpu.set_input(mode, cmd)
    orc.make_step(0.01, 1000, n_threads=os.cpu_count())
    for _ in range(1000):
        gpu.make_step(0.01)
    so, sg = orc.get_state(), gpu.get_full_state()
    ex = np.max(np.abs(so["x"] - sg["x"]), axis=1)
    rel = ex / (1.0 + np.max(np.abs(so["x"])))
    out[name] = {"x_abs_max": float(ex.max()), "x_abs_p50": float(np.median(ex)), "x_abs_p99": float(np.quantile(ex, 0.99)),
                 "x_rel_to_excursion_max": float(rel.max()), "R_abs_max": float(np.max(np.abs(so["R"] - sg["R"]))),
                 "uavs_above_1e-7_rel": int((rel > 1e-7).sum()), "excursion": float(np.max(np.abs(so["x"])))}
print(json.dumps(out))
