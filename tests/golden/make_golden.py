#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/.

The reference ships no golden vectors, so they are generated here from the reference itself:

* trajectory fixtures (c1_position_x500.json, modes_2s.json) come from the reference's OWN
  UavSystem sources compiled in this container (oracle/_ref/libref_uavsystem.so — uav_system.hpp,
  multirotor_model.hpp and controllers/*.hpp from /root/reference/include, against the Eigen/odeint
  stand-ins of oracle/shim because the image has neither library; see oracle/ref_uavsystem.cpp).
  The generator also checks that the restated oracle reproduces every number bit for bit;
* the collision fixture comes from the reference's REAL vendored nanoflann
  (oracle/_ref/libref_nanoflann.so, built from /root/reference/include).

    python tests/golden/make_golden.py        # rewrites tests/golden/*.json
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import grid_spawn, rand  # noqa: E402
from mrs_multirotor_simulator_b200.airframes import airframe  # noqa: E402
from oracle import binding as O  # noqa: E402


def git_hash():
    try:
        return subprocess.check_output(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], text=True).strip()
    except Exception:
        return "unknown"


def state_dict(st, i=0):
    return {k: [float(v) for v in st[k][i]] for k in ("x", "v", "R", "omega", "motor_rpm", "imu")}


def c1():
    """BASELINE config 1: x500, spawn (0,0,1) heading 0, PositionCmd (5,-3,4) heading 1.0, dt=0.005, 10 s."""
    s, chk = (cls([airframe("x500")], spawn_xyz=[[0, 0, 1]], spawn_heading=[0.0], n=1) for cls in (O.RefSwarm, O.OracleSwarm))
    samples = []
    for u in (s, chk):
        u.set_input(O.POSITION_CMD, [[5.0, -3.0, 4.0, 1.0]])
    for sec in range(10):
        s.make_step(0.005, 200)
        chk.make_step(0.005, 200)
        assert state_dict(s.get_state()) == state_dict(chk.get_state()), "restated oracle differs from the compiled reference"
        samples.append({"t": sec + 1, **state_dict(s.get_state())})
    return {"config": "C1", "frame": "x500", "spawn": [0, 0, 1], "heading": 0.0, "cmd": [5.0, -3.0, 4.0, 1.0], "dt": 0.005, "samples": samples}


MODE_CMDS = {
    "ACTUATOR_CMD": (O.ACTUATOR_CMD, [0.55, 0.56, 0.57, 0.58, 0, 0, 0, 0]),
    "CONTROL_GROUP_CMD": (O.CONTROL_GROUP_CMD, [0.02, -0.01, 0.03, 0.55]),
    "ATTITUDE_RATE_CMD": (O.ATTITUDE_RATE_CMD, [0.1, -0.2, 0.3, 0.55]),
    "ATTITUDE_CMD": (O.ATTITUDE_CMD, [0.9553364891256060, 0.2955202066613396, 0.0, -0.2896294776255156, 0.9362933635841992, 0.1986693307950612,
                                      0.0587108016938265, -0.1897961087496159, 0.9800665778412416, 0.56]),
    "TILT_HDG_RATE_CMD": (O.TILT_HDG_RATE_CMD, [0.1, -0.15, 1.0, 0.4, 0.56]),
    "ACCELERATION_HDG_RATE_CMD": (O.ACCELERATION_HDG_RATE_CMD, [0.5, -1.0, 0.3, 0.5]),
    "ACCELERATION_HDG_CMD": (O.ACCELERATION_HDG_CMD, [0.5, -1.0, 0.3, -2.0]),
    "VELOCITY_HDG_RATE_CMD": (O.VELOCITY_HDG_RATE_CMD, [1.5, -1.0, 0.8, 0.5]),
    "VELOCITY_HDG_CMD": (O.VELOCITY_HDG_CMD, [1.5, -1.0, 0.8, 2.5]),
    "POSITION_CMD": (O.POSITION_CMD, [3.0, 2.0, 7.0, -1.0]),
}


def modes():
    """Every input mode on every motor count: 2 s (200 steps of 0.01) from spawn (1,2,5) heading 0.3."""
    out = []
    for frame in ("x500", "f550", "naki"):
        for name, (mode, cmd) in MODE_CMDS.items():
            s, chk = (cls([airframe(frame)], spawn_xyz=[[1, 2, 5]], spawn_heading=[0.3], n=1) for cls in (O.RefSwarm, O.OracleSwarm))
            for u in (s, chk):
                u.set_input(mode, [cmd])
                u.make_step(0.01, 200)
            assert state_dict(s.get_state()) == state_dict(chk.get_state()), "restated oracle differs from the compiled reference"
            out.append({"frame": frame, "mode": name, "mode_id": mode, "cmd": cmd, "steps": 200, "dt": 0.01, **state_dict(s.get_state())})
    return {"spawn": [1, 2, 5], "heading": 0.3, "cases": out}


def collisions():
    """400 mixed-type UAVs on a jittered 1.6 m grid: directed pair list and forces from the REAL nanoflann."""
    n = 400
    frames = ["x500", "f550", "naki", "t650"]
    tou = (np.arange(n) * 3 % 4).astype(int)
    k = np.arange(n)
    xyz = np.stack([1.6 * (k % 20) + rand(9, 0, n, -0.7, 0.7), 1.6 * (k // 20) + rand(9, 1, n, -0.7, 0.7), rand(9, 2, n, 2.0, 3.2)], axis=1)
    arm = np.array([airframe(f)["arm_length"] for f in frames])[tou]
    prop = np.array([airframe(f)["prop_radius"] for f in frames])[tou]
    mass = np.array([airframe(f)["mass"] for f in frames])[tou]
    pairs, forces, _ = O.collide_snapshot(xyz, arm, prop, mass, False, 100.0, engine="nanoflann")
    _, _, crashed = O.collide_snapshot(xyz, arm, prop, mass, True, 100.0, engine="nanoflann")
    pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
    return {"engine": "reference nanoflann v1.5.0 (include/nanoflann.hpp), loop of src/multirotor_simulator.cpp:303-358", "frames": frames,
            "type_of_uav": tou.tolist(), "xyz": [[float(v) for v in row] for row in xyz], "rebounce": 100.0, "pairs": pairs.tolist(),
            "forces": [[float(v) for v in row] for row in forces], "crashed": crashed.astype(int).tolist()}


if __name__ == "__main__":
    O.build()
    assert O.ref_lib() is not None, "oracle/_ref/libref_nanoflann.so is needed (build it where /root/reference exists)"
    assert O.refsys_lib() is not None, "oracle/_ref/libref_uavsystem.so is needed (build it where /root/reference exists)"
    meta = {"generator": "tests/golden/make_golden.py", "oracle_git": git_hash(),
            "trajectories_from": "reference UavSystem sources compiled against oracle/shim (oracle/_ref/libref_uavsystem.so)",
            "collisions_from": "reference nanoflann (oracle/_ref/libref_nanoflann.so)"}
    for name, fn in (("c1_position_x500", c1), ("modes_2s", modes), ("collisions_400", collisions)):
        doc = {"meta": meta, **fn()}
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(doc, f, indent=0)
        print("wrote", name)
