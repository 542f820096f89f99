"""The CUDA stepping path against the reference's OWN UavSystem sources (oracle/_ref/libref_uavsystem.so:
uav_system.hpp, multirotor_model.hpp, controllers/*.hpp compiled unmodified against the Eigen/odeint
stand-ins of oracle/shim — built in the container that has /root/reference, shipped prebuilt to the GPU
box).  Same tolerances as against the restated oracle (helpers.TOL)."""
import numpy as np
import pytest

from helpers import TOL, grid_spawn, rand
from oracle import binding as O
from test_step_parity import ALL_MODES, _commands, af

pytestmark = pytest.mark.gpu

if O.refsys_lib() is None:
    pytest.skip("oracle/_ref/libref_uavsystem.so was not shipped", allow_module_level=True)


def make(types, tou, spawn, heading=None):
    from mrs_multirotor_simulator_b200 import UavBatch

    n = len(spawn)
    heading = np.zeros(n) if heading is None else heading
    ref = O.RefSwarm(types, type_of_uav=tou, spawn_xyz=spawn, spawn_heading=heading, n=n)
    gpu = UavBatch(types, type_of_uav=tou, spawn_xyz=spawn, spawn_heading=heading, n=n, device=0)
    return ref, gpu


def check(ref, gpu, what, tol=TOL):
    sr, sg = ref.get_state(), gpu.get_full_state()
    bad = []
    for k, t in tol.items():
        assert np.all(np.isfinite(sr[k])), f"reference {k} not finite ({what})"
        d = float(np.max(np.abs(sr[k] - sg[k])))
        if not d <= t:
            bad.append(f"{k}: {d:.3e} > {t:.1e}")
    assert not bad, f"{what}: " + "; ".join(bad)


def test_c1_single_x500_position_cmd_10s_vs_compiled_reference():
    ref, gpu = make([af("x500")], None, np.array([[0.0, 0.0, 1.0]]))
    for s in (ref, gpu):
        s.set_input(O.POSITION_CMD, [[5.0, -3.0, 4.0, 1.0]])
    ref.make_step(0.005, 2000)
    for _ in range(2000):
        gpu.make_step(0.005)
    check(ref, gpu, "C1")


@pytest.mark.parametrize("frame", ["x500", "f550", "naki"])
@pytest.mark.parametrize("mode", [m for m in ALL_MODES if m not in (O.ACTUATOR_CMD, O.CONTROL_GROUP_CMD, O.ATTITUDE_RATE_CMD)])
def test_10s_closed_loop_modes_vs_compiled_reference(mode, frame):
    n = 64
    ref, gpu = make([af(frame)], None, grid_spawn(n, z=10.0), rand(3, 0, n, -3, 3))
    cmd = _commands(mode, n)
    for s in (ref, gpu):
        s.set_input(mode, cmd)
    ref.make_step(0.01, 1000, n_threads=4)
    if mode == O.VELOCITY_HDG_CMD:  # the K-fused launch too
        for _ in range(100):
            gpu.make_step(0.01, 10)
    else:
        for _ in range(1000):
            gpu.make_step(0.01)
    check(ref, gpu, f"mode {mode} {frame}")


@pytest.mark.parametrize("mode", [O.ACTUATOR_CMD, O.CONTROL_GROUP_CMD, O.ATTITUDE_RATE_CMD])
def test_open_loop_modes_vs_compiled_reference(mode):
    """Open loop in attitude: compare one second tightly, before the tumbling amplifies rounding."""
    n = 64
    ref, gpu = make([af("x500")], None, grid_spawn(n, z=10.0), rand(3, 0, n, -3, 3))
    cmd = _commands(mode, n)
    for s in (ref, gpu):
        s.set_input(mode, cmd)
    ref.make_step(0.01, 100)
    for _ in range(100):
        gpu.make_step(0.01)
    check(ref, gpu, f"mode {mode}")


def test_c5_mixed_airframes_and_events_vs_compiled_reference():
    """Mixed 4/6/8-motor swarm, ground plane, custom gains, feed-forward, crash, force — 6 s."""
    n = 96
    types = [af(f, ground_enabled=True, ground_z=0.0) for f in ("x500", "f550", "naki")]
    tou = (np.arange(n) % 3).astype(np.int32)
    ref, gpu = make(types, tou, grid_spawn(n, z=0.0))
    cmd = _commands(O.POSITION_CMD, n)
    for s in (ref, gpu):
        s.set_input(O.POSITION_CMD, cmd)
        s.set_controller_params("position", [1.5, 0.1, 0.1, 3.0], [0, 1, 2, 3])
    ref.set_feedforward(0, np.tile([0.2, 0.1, -0.1, 0.4], (n, 1)))
    gpu.set_feedforward("acceleration_hdg_rate", np.tile([0.2, 0.1, -0.1, 0.4], (n, 1)))
    ref.make_step(0.01, 300)
    for _ in range(300):
        gpu.make_step(0.01)
    check(ref, gpu, "climb")
    for s in (ref, gpu):
        s.crash([5, 6, 7])
        s.apply_force(np.tile([1.0, -2.0, 0.5], (n, 1)))
    ref.make_step(0.01, 300)
    for _ in range(300):
        gpu.make_step(0.01)
    check(ref, gpu, "crash + force")
    assert np.all(gpu.get_state()["x"][[5, 6, 7], 2] == 0.0)  # the crashed ones fell onto the ground plane
