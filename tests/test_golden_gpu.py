"""GPU path vs the committed golden fixtures (tests/golden/, see make_golden.py for provenance)."""
import json
import os

import numpy as np
import pytest

from helpers import TOL

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def check(st, ref, what):
    for k in ("x", "v", "R", "omega", "motor_rpm", "imu"):
        d = np.max(np.abs(st[k][0] - np.array(ref[k])))
        assert d <= TOL[k], f"{what} {k}: {d:.3e}"


def test_c1_trajectory_against_golden():
    from mrs_multirotor_simulator_b200 import POSITION_CMD, UavBatch, airframe

    g = load("c1_position_x500.json")
    b = UavBatch([airframe("x500")], spawn_xyz=[g["spawn"]], spawn_heading=[g["heading"]], n=1)
    b.set_input(POSITION_CMD, [g["cmd"]])
    for smp in g["samples"]:
        for _ in range(200):
            b.make_step(g["dt"])
        check(b.get_full_state(), smp, f"t={smp['t']}")


def test_every_mode_every_motor_count_against_golden():
    from mrs_multirotor_simulator_b200 import UavBatch, airframe

    g = load("modes_2s.json")
    for c in g["cases"]:
        b = UavBatch([airframe(c["frame"])], spawn_xyz=[g["spawn"]], spawn_heading=[g["heading"]], n=1)
        b.set_input(c["mode_id"], [c["cmd"]])
        for _ in range(c["steps"] // 10):
            b.make_step(c["dt"], 10)
        check(b.get_full_state(), c, f"{c['frame']} {c['mode']}")


def test_reference_generated_collision_fixture():
    """Pair list, forces and crash flags produced by the reference's real nanoflann."""
    from mrs_multirotor_simulator_b200 import UavBatch, airframe

    g = load("collisions_400.json")
    types = [airframe(f) for f in g["frames"]]
    tou = np.array(g["type_of_uav"], dtype=np.int32)
    n = len(tou)
    b = UavBatch(types, type_of_uav=tou, spawn_xyz=np.zeros((n, 3)), n=n)
    b.set_state(x=np.array(g["xyz"]))  # positions arrive through set_state, not through spawn
    b.set_collisions(True, False, g["rebounce"])
    b.handle_collisions()
    assert b.get_collision_pairs().tolist() == g["pairs"]
    assert np.allclose(b.get_force(), np.array(g["forces"]), rtol=1e-12, atol=0)
    assert not b.has_crashed().any()
    b.set_collisions(True, True, g["rebounce"])
    b.handle_collisions()
    assert b.has_crashed().tolist() == g["crashed"]
    assert not b.get_force().any()
