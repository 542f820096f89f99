"""Scenario loader (mrs_multirotor_simulator_b200/scenario.py): the reference's YAML configuration
files -> the swarm the reference node would have built."""
import glob
import os

import numpy as np
import pytest

from mrs_multirotor_simulator_b200 import AIRFRAMES, CONTROLLER_DEFAULTS
from mrs_multirotor_simulator_b200.scenario import load_scenario, scenario_from_tree

REF = "/root/reference"

CUSTOM = """
collisions:
  crash: false
  rebounce: 50.0
ground:
  z: -1.5
individual_takeoff_platform:
  enabled: true
simulation_rate: 250.0
velocity_controller:
  kp: 3.0
mixer:
  desaturation: false
uav_names: ["a", "b", "c"]
a:
  type: "x500"
  spawn: {x: 1.0, y: 2.0, z: 3.0, heading: 0.5}
b:
  type: "mini"
  spawn: {x: -4.0, y: 0.0, z: 0.0, heading: 0}
c:
  type: "x500"
  spawn: {x: 0.0, y: 8.0, z: 1.0, heading: -1.0}
mini:
  n_motors: 4
  mass: 0.5
  arm_length: 0.1
  body_height: 0.04
  motor_time_constant: 0.02
  air_resistance_coeff: 0.2
  propulsion:
    prop_radius: 0.05
    force_constant: 0.00000001
    moment_constant: 0.01
    rpm: {min: 2000, max: 30000}
    allocation_matrix: [-0.707, 0.707, 0.707, -0.707,
                        -0.707, 0.707, -0.707, 0.707,
                        -1, -1, 1, 1,
                        1, 1, 1, 1]
"""


def test_custom_config_alone_overlays_the_shipped_defaults(tmp_path):
    p = tmp_path / "simulator.yaml"
    p.write_text(CUSTOM)
    s = load_scenario(str(p))
    assert s.uav_names == ["a", "b", "c"] and s.type_names == ["x500", "mini"] and list(s.type_of_uav) == [0, 1, 0]
    assert np.array_equal(s.spawn_xyz, [[1, 2, 3], [-4, 0, 0], [0, 8, 1]]) and np.array_equal(s.spawn_heading, [0.5, 0.0, -1.0])
    assert s.dt == 1 / 250.0 and s.collisions_enabled and not s.collisions_crash and s.collisions_rebounce == 50.0
    x500, mini = s.types
    assert {k: x500[k] for k in AIRFRAMES["x500"]} == AIRFRAMES["x500"]
    assert mini["mass"] == 0.5 and mini["n_motors"] == 4 and mini["kf"] == 1e-8 and mini["min_rpm"] == 2000.0 and len(mini["allocation"]) == 4
    for t in s.types:  # world parameters reach every airframe (uav_system_ros.cpp:51-56)
        assert t["g"] == 9.81 and t["ground_enabled"] and t["ground_z"] == -1.5 and t["takeoff_patch_enabled"]
    want = dict(CONTROLLER_DEFAULTS, vel_kp=3.0, mixer_desaturation=False)
    assert s.controllers == want
    assert s.input_timeout == 1.0 and s.iterate_without_input


def test_later_files_win_key_by_key(tmp_path):
    a, b = tmp_path / "a.yaml", tmp_path / "b.yaml"
    a.write_text("collisions: {enabled: true, crash: true, rebounce: 100.0}\nuav_names: [u]\nu: {type: f550, spawn: {x: 0, y: 0, z: 0, heading: 0}}\n")
    b.write_text("collisions: {crash: false}\nu: {spawn: {z: 2.5}}\n")
    s = load_scenario(str(a), str(b))
    assert s.collisions_enabled and not s.collisions_crash and s.collisions_rebounce == 100.0
    assert list(s.spawn_xyz[0]) == [0.0, 0.0, 2.5] and s.type_names == ["f550"]


def test_errors_are_reported_like_a_failed_param_load():
    with pytest.raises(ValueError):
        scenario_from_tree({"uav_names": []})
    with pytest.raises(ValueError):
        scenario_from_tree({"uav_names": ["u"]})
    with pytest.raises(ValueError):
        scenario_from_tree({"uav_names": ["u"], "u": {"type": "nonsense", "spawn": {"x": 0, "y": 0, "z": 0, "heading": 0}}})


def test_spawn_randomisation_is_bounded_and_reproducible():
    tree = {"uav_names": [f"u{i}" for i in range(50)], "randomization": {"enabled": True, "bounds": {"x": 2.0, "y": 3.0, "z": 0.5}}}
    for i in range(50):
        tree[f"u{i}"] = {"type": "x500", "spawn": {"x": 10.0 * i, "y": 0.0, "z": 5.0, "heading": 0.0}}
    a, b, c = scenario_from_tree(tree, seed=1), scenario_from_tree(tree, seed=1), scenario_from_tree(tree, seed=2)
    assert np.array_equal(a.spawn_xyz, b.spawn_xyz) and not np.array_equal(a.spawn_xyz, c.spawn_xyz)
    d = a.spawn_xyz - np.stack([10.0 * np.arange(50), np.zeros(50), np.full(50, 5.0)], axis=1)
    assert np.all(np.abs(d) <= [2.0, 3.0, 0.5]) and np.abs(d).max() > 0.3
    assert np.all(np.abs(a.spawn_heading) <= 3.14) and np.ptp(a.spawn_heading) > 1.0


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "config")), reason="reference tree not present")
def test_the_reference_s_own_files_run_unmodified():
    """config/*.yaml + config/uavs/*.yaml + config/controllers/*.yaml + the 400-UAV scenario's custom config."""
    files = ([os.path.join(REF, "config", "multirotor_simulator.yaml"), os.path.join(REF, "config", "uavs.yaml")]
             + sorted(glob.glob(os.path.join(REF, "config", "uavs", "*.yaml"))) + sorted(glob.glob(os.path.join(REF, "config", "controllers", "*.yaml"))))
    one = load_scenario(*files)
    assert one.uav_names == ["uav1"] and one.type_names == ["x500"] and list(one.spawn_xyz[0]) == [10.0, 15.0, 0.0] and one.spawn_heading[0] == 3.14
    assert one.collisions_enabled and one.collisions_crash and one.collisions_rebounce == 100.0 and one.dt == 0.01
    assert one.controllers == CONTROLLER_DEFAULTS  # the YAML defaults equal the header defaults
    assert {k: one.types[0][k] for k in AIRFRAMES["x500"]} == AIRFRAMES["x500"]  # the shipped table is the content of config/uavs/x500.yaml
    assert one.types[0]["ground_enabled"] and one.types[0]["ground_z"] == 0.0 and not one.types[0]["takeoff_patch_enabled"]
    assert one.frames["world"]["name"] == "simulator_origin"

    swarm = load_scenario(*files, os.path.join(REF, "tmux", "standalone_400_uavs", "custom_configs", "simulator.yaml"))
    assert swarm.n == 400 and swarm.type_names == ["f550"] and not swarm.collisions_crash and swarm.collisions_rebounce == 100.0
    # the 20 x 20 grid, 4 m pitch (the file lists it as four 10 x 10 blocks)
    assert {tuple(p) for p in swarm.spawn_xyz} == {(4.0 * i, 4.0 * j, 0.0) for i in range(20) for j in range(20)}
    assert list(swarm.spawn_xyz[10]) == [0.0, 4.0, 0.0] and swarm.uav_names[10] == "uav11"
    assert not swarm.spawn_heading.any()
    assert {k_: swarm.types[0][k_] for k_ in AIRFRAMES["f550"]} == AIRFRAMES["f550"]
    # every shipped airframe table entry is the content of the reference's file
    for path in glob.glob(os.path.join(REF, "config", "uavs", "*.yaml")):
        name = os.path.splitext(os.path.basename(path))[0]
        import yaml

        with open(path) as f:
            s = scenario_from_tree({**yaml.safe_load(f), "uav_names": ["u"], "u": {"type": name, "spawn": {"x": 0, "y": 0, "z": 0, "heading": 0}}})
        assert {k_: s.types[0][k_] for k_ in AIRFRAMES[name]} == AIRFRAMES[name], name


@pytest.mark.gpu
def test_scenario_batch_equals_the_hand_built_swarm(tmp_path):
    """make_batch == the C2 construction used elsewhere in the tests, bit for bit after 200 ticks."""
    from helpers import grid_spawn, rand
    from mrs_multirotor_simulator_b200 import ACTUATOR_CMD, VELOCITY_HDG_RATE_CMD, UavBatch, airframe

    n = 100
    lines = ["collisions: {crash: false}", "uav_names: [" + ", ".join(f"uav{i + 1}" for i in range(n)) + "]"]
    spawn = grid_spawn(n, pitch=4.0, z=0.0)
    for i in range(n):
        lines.append(f"uav{i + 1}: {{type: f550, spawn: {{x: {spawn[i, 0]}, y: {spawn[i, 1]}, z: 0.0, heading: 0}}}}")
    p = tmp_path / "simulator.yaml"
    p.write_text("\n".join(lines) + "\n")
    a = load_scenario(str(p)).make_batch()
    b = UavBatch([airframe("f550", ground_enabled=True, ground_z=0.0)], spawn_xyz=spawn, n=n)
    b.set_input(ACTUATOR_CMD, np.zeros((n, 8)))
    b.make_step(0.01)
    b.make_step(0.01)
    b.set_collisions(True, False, 100.0)
    cmd = np.stack([rand(42, 1, n, -2, 2), rand(42, 2, n, -2, 2), rand(42, 3, n, 0, 2), rand(42, 4, n, -1, 1)], axis=1)
    for s in (a, b):
        s.set_input(VELOCITY_HDG_RATE_CMD, cmd)
        s.run(0.01, 200, with_collisions=True)
    sa, sb = a.get_full_state(), b.get_full_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
