"""The C++ façade (include/mrsb/uav_system.hpp): compiles and links against libmrsb.so on any
machine; on a GPU it reproduces BASELINE config 1 through the reference's own method names."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "hover_to_waypoint.cpp")
PKG = os.path.join(ROOT, "mrs_multirotor_simulator_b200")


def build(tmp_path, src=SRC):
    exe = str(tmp_path / os.path.splitext(os.path.basename(src))[0])
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", PKG, "-lmrsb",
                    f"-Wl,-rpath,{PKG}"], check=True)
    return exe


NODE = os.path.join(ROOT, "tests", "cpp", "node_loop.cpp")


def test_facade_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    import torch

    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_facade_reproduces_config_1(tmp_path):
    exe = build(tmp_path)
    out = json.loads(subprocess.check_output([exe], text=True))
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "c1_position_x500.json")))["samples"][-1]
    assert np.max(np.abs(np.array(out["x"]) - gold["x"])) <= 1e-9
    assert np.max(np.abs(np.array(out["v"]) - gold["v"])) <= 1e-9
    assert np.max(np.abs(np.array(out["rpm"]) - gold["motor_rpm"][:4])) <= 1e-6
    assert out["n_motors"] == 4 and out["crashed"] == [0, 1, 1] and out["pairs"] == 2


def test_node_loop_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    import torch

    exe = build(tmp_path, NODE)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_node_loop_through_the_facade_equals_the_python_mirror(tmp_path):
    """The reference node's loop (makeStep for all + handleCollisions, iterate_without_input off) driven from C++ through
    Swarm::run equals the same flight driven through the Python mirror, bit for bit; the UAVs without a command never move."""
    from mrs_multirotor_simulator_b200 import VELOCITY_HDG_RATE_CMD, UavBatch, airframe

    out = json.loads(subprocess.check_output([build(tmp_path, NODE)], text=True))
    n = 256
    i = np.arange(n)
    spawn = np.stack([2.0 * (i % 16), 2.0 * (i // 16), np.full(n, 3.0)], axis=1)
    b = UavBatch([airframe("x500", ground_enabled=True, ground_z=0.0, takeoff_patch_enabled=False)], spawn_xyz=spawn, n=n)
    b.set_iterate_without_input(False)
    b.set_collisions(True, False, 100.0)
    even = np.arange(0, n, 2, dtype=np.int32)
    cmd = np.stack([0.5 * ((even % 7) - 3), 0.4 * ((even % 5) - 2), 0.1 * (even % 3), np.full(len(even), 0.2)], axis=1)
    for k, e in enumerate(even):  # one setInput per UAV, like the C++ program
        b.set_input(VELOCITY_HDG_RATE_CMD, cmd[k:k + 1], idx=np.array([e], dtype=np.int32))
    b.run(0.01, 200)
    st = b.get_state()
    s = 0.0
    for k in range(n):
        s += st["x"][k, 0] + 2.0 * st["x"][k, 1] + 3.0 * st["x"][k, 2]
    assert out["sum"] == s
    assert out["idle_z"] == 3.0 and out["idle_rpm0"] == 0.0
    assert out["pairs"] == len(b.get_collision_pairs())
