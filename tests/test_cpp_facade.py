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


def build(tmp_path):
    exe = str(tmp_path / "hover_to_waypoint")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", PKG, "-lmrsb",
                    f"-Wl,-rpath,{PKG}"], check=True)
    return exe


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
