"""GPU stepping kernel vs the CPU oracle (UavSystem::makeStep, uav_system.hpp:304-380)."""
import numpy as np
import pytest

from helpers import TOL, assert_parity, grid_spawn, make_pair, max_diffs, rand
from oracle import binding as O

pytestmark = pytest.mark.gpu


def af(name, **kw):
    from mrs_multirotor_simulator_b200 import airframe

    return airframe(name, **kw)


def test_c1_single_x500_position_cmd_10s():
    """BASELINE config 1: one x500, PositionCmd hover-to-waypoint, RK4 dt=0.005, 10 s."""
    orc, gpu = make_pair([af("x500")], None, np.array([[0.0, 0.0, 1.0]]))
    cmd = [[5.0, -3.0, 4.0, 1.0]]
    orc.set_input(O.POSITION_CMD, cmd)
    gpu.set_input(O.POSITION_CMD, cmd)
    for _ in range(20):
        orc.make_step(0.005, 100)
        for _ in range(100):
            gpu.make_step(0.005)
    assert_parity(orc, gpu, what="C1")
    x = gpu.get_state()["x"][0]
    # the independent numpy probe of SURVEY App. D ends at (4.99966878, -2.99316222, 3.99492282)
    assert np.allclose(x, [4.99966878, -2.99316222, 3.99492282], atol=2e-8)


def _commands(mode, n, seed=42):
    r = lambda s, lo, hi: rand(seed, s, n, lo, hi)
    if mode == O.ACTUATOR_CMD:
        return np.stack([rand(seed, 10 + m, n, 0.4, 0.7) for m in range(8)], axis=1)
    if mode == O.CONTROL_GROUP_CMD:
        return np.stack([r(1, -0.05, 0.05), r(2, -0.05, 0.05), r(3, -0.05, 0.05), r(4, 0.45, 0.65)], axis=1)
    if mode == O.ATTITUDE_RATE_CMD:
        return np.stack([r(1, -0.3, 0.3), r(2, -0.3, 0.3), r(3, -0.5, 0.5), r(4, 0.45, 0.65)], axis=1)
    if mode == O.ATTITUDE_CMD:
        roll, pitch, yaw = r(1, -0.3, 0.3), r(2, -0.3, 0.3), r(3, -3.0, 3.0)
        out = np.zeros((n, 10))
        for i in range(n):
            cr, sr, cp, sp, cy, sy = np.cos(roll[i]), np.sin(roll[i]), np.cos(pitch[i]), np.sin(pitch[i]), np.cos(yaw[i]), np.sin(yaw[i])
            Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
            Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
            Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
            out[i, :9] = (Rz @ Ry @ Rx).T.reshape(9)  # column-major
        out[:, 9] = r(4, 0.45, 0.65)
        return out
    if mode == O.TILT_HDG_RATE_CMD:
        return np.stack([r(1, -0.3, 0.3), r(2, -0.3, 0.3), r(3, 0.8, 1.2), r(4, -1.0, 1.0), r(5, 0.45, 0.65)], axis=1)
    if mode in (O.ACCELERATION_HDG_RATE_CMD, O.ACCELERATION_HDG_CMD):
        return np.stack([r(1, -2, 2), r(2, -2, 2), r(3, -1, 2), r(4, -1.0, 1.0) if mode == O.ACCELERATION_HDG_RATE_CMD else r(4, -3.1, 3.1)], axis=1)
    if mode in (O.VELOCITY_HDG_RATE_CMD, O.VELOCITY_HDG_CMD):
        return np.stack([r(1, -2, 2), r(2, -2, 2), r(3, 0, 2), r(4, -1.0, 1.0) if mode == O.VELOCITY_HDG_RATE_CMD else r(4, -3.1, 3.1)], axis=1)
    if mode == O.POSITION_CMD:
        return np.stack([r(1, -10, 10), r(2, -10, 10), r(3, 2, 12), r(4, -3.1, 3.1)], axis=1)
    raise ValueError(mode)


ALL_MODES = [O.ACTUATOR_CMD, O.CONTROL_GROUP_CMD, O.ATTITUDE_RATE_CMD, O.ATTITUDE_CMD, O.TILT_HDG_RATE_CMD, O.ACCELERATION_HDG_RATE_CMD,
             O.ACCELERATION_HDG_CMD, O.VELOCITY_HDG_RATE_CMD, O.VELOCITY_HDG_CMD, O.POSITION_CMD]


@pytest.mark.parametrize("mode", ALL_MODES)
@pytest.mark.parametrize("frame", ["x500", "f550", "naki"])
def test_single_step_every_mode(mode, frame):
    """One makeStep from a perturbed state: rounding-level agreement (catches any semantic slip)."""
    n = 64
    spawn = grid_spawn(n, z=5.0)
    orc, gpu = make_pair([af(frame)], None, spawn, rand(7, 0, n, -3, 3))
    # perturbed state: moving, tilted, spinning motors
    st = orc.get_state()
    v = np.stack([rand(7, 1, n, -2, 2), rand(7, 2, n, -2, 2), rand(7, 3, n, -1, 1)], axis=1)
    w = np.stack([rand(7, 4, n, -0.5, 0.5), rand(7, 5, n, -0.5, 0.5), rand(7, 6, n, -0.5, 0.5)], axis=1)
    rpm = np.zeros((n, 8))
    nm = af(frame)["n_motors"]
    rpm[:, :nm] = np.stack([rand(7, 20 + m, n, 3000, 5000) for m in range(nm)], axis=1)
    for s in (orc, gpu):
        s.set_state(x=st["x"], v=v, R=st["R"], omega=w, motor_rpm=rpm)
    cmd = _commands(mode, n)
    orc.set_input(mode, cmd)
    gpu.set_input(mode, cmd)
    for step in range(3):
        orc.make_step(0.01)
        gpu.make_step(0.01)
        tight = {"x": 1e-12, "v": 1e-11, "R": 1e-13, "omega": 1e-9, "motor_rpm": 1e-8, "v_prev": 1e-11, "imu": 1e-8}
        assert_parity(orc, gpu, tol=tight, what=f"mode {mode} {frame} step {step}")


@pytest.mark.parametrize("mode", ALL_MODES)
def test_10s_every_mode_x500(mode):
    """1000 steps of 0.01 s per input mode on 64 x500s with seeded random references."""
    n = 64
    orc, gpu = make_pair([af("x500")], None, grid_spawn(n, z=10.0), rand(3, 0, n, -3, 3))
    cmd = _commands(mode, n)
    orc.set_input(mode, cmd)
    gpu.set_input(mode, cmd)
    orc.make_step(0.01, 1000)
    for _ in range(1000):
        gpu.make_step(0.01)
    if mode in (O.ACTUATOR_CMD, O.CONTROL_GROUP_CMD, O.ATTITUDE_RATE_CMD):
        # open-loop in position/attitude: the flight is unstable (tumbling or km-scale drift); errors
        # grow with the state itself, so compare relative to its magnitude
        so, sg = orc.get_state(), gpu.get_full_state()
        for k in ("x", "v"):
            scale = 1.0 + np.max(np.abs(so[k]))
            assert np.max(np.abs(so[k] - sg[k])) <= 1e-7 * scale, k
        assert np.max(np.abs(so["motor_rpm"] - sg["motor_rpm"])) <= 1e-5
    else:
        assert_parity(orc, gpu, what=f"mode {mode}")


def test_k_fused_substeps_equal_separate_launches():
    """mrsb_make_step(dt, K) == K x mrsb_make_step(dt, 1), bit for bit."""
    n = 256
    from mrs_multirotor_simulator_b200 import UavBatch

    cmd = _commands(O.VELOCITY_HDG_CMD, n)
    res = []
    for k in (1, 10):
        b = UavBatch([af("x500")], spawn_xyz=grid_spawn(n, z=10.0), n=n)
        b.set_input(O.VELOCITY_HDG_CMD, cmd)
        for _ in range(100 // k):
            b.make_step(0.01, k)
        res.append(b.get_full_state())
    for key in res[0]:
        assert np.array_equal(res[0][key], res[1][key]), key


def test_c3_rl_batch_velocity_hdg_k10():
    """BASELINE config 3 at reduced N: VelocityHdgCmd, K=10 fused substeps, 100 launches = 10 s."""
    n = 1024
    orc, gpu = make_pair([af("x500")], None, grid_spawn(n, z=10.0))
    cmd = np.stack([rand(42, 1, n, -2, 2), rand(42, 2, n, -2, 2), rand(42, 3, n, -2, 2), rand(42, 4, n, -np.pi, np.pi)], axis=1)
    orc.set_input(O.VELOCITY_HDG_CMD, cmd)
    gpu.set_input(O.VELOCITY_HDG_CMD, cmd)
    orc.make_step(0.01, 1000, n_threads=8)
    for _ in range(100):
        gpu.make_step(0.01, 10)
    assert_parity(orc, gpu, what="C3")


def test_c5_mixed_airframes_actuator_cmd():
    """BASELINE config 5 at reduced N: x500/f550/naki interleaved, open-loop motors redrawn every 100 steps, ground on."""
    n = 768
    types = [af(f, ground_enabled=True, ground_z=0.0) for f in ("x500", "f550", "naki")]
    tou = (np.arange(n) % 3).astype(np.int32)
    orc, gpu = make_pair(types, tou, grid_spawn(n, z=0.0))
    for epoch in range(5):
        cmd = np.stack([rand(42 + epoch, 10 + m, n, 0.4, 0.7) for m in range(8)], axis=1)
        orc.set_input(O.ACTUATOR_CMD, cmd)
        gpu.set_input(O.ACTUATOR_CMD, cmd)
        orc.make_step(0.01, 100, n_threads=8)
        for _ in range(100):
            gpu.make_step(0.01)
    so, sg = orc.get_state(), gpu.get_full_state()
    assert np.max(np.abs(so["motor_rpm"] - sg["motor_rpm"])) <= 1e-6
    for k in ("x", "v"):  # open loop: unstable attitude, compare relative to the excursion
        scale = 1.0 + np.max(np.abs(so[k]))
        assert np.max(np.abs(so[k] - sg[k])) <= 1e-7 * scale, k


def test_ground_and_takeoff_patch():
    n = 32
    ground = af("x500", ground_enabled=True, ground_z=0.0)
    patch = af("f550", takeoff_patch_enabled=True)
    tou = (np.arange(n) % 2).astype(np.int32)
    spawn = grid_spawn(n, z=0.0)
    spawn[1::2, 2] = 3.0  # the patch holds a UAV at its spawn height until the motors spin up
    orc, gpu = make_pair([ground, patch], tou, spawn)
    # idle: everybody stays put (ground clamp / take-off platform)
    orc.make_step(0.01, 50)
    for _ in range(50):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="idle")
    assert np.allclose(gpu.get_state()["x"][:, 2], spawn[:, 2])
    cmd = np.tile([0.0, 0.0, 1.0, 0.3], (n, 1))
    orc.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
    gpu.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
    orc.make_step(0.01, 500)
    for _ in range(500):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="takeoff")
    assert np.all(gpu.get_state()["x"][:, 2] > spawn[:, 2] + 1.0)
    assert gpu.get_params(1).takeoff_patch_enabled == 0 and orc.get_params(1).takeoff_patch_enabled == 0
    # land on the ground again
    cmd = np.tile([0.0, 0.0, -1.5, 0.0], (n, 1))
    orc.set_input(O.VELOCITY_HDG_RATE_CMD, cmd[0::2], idx=np.arange(0, n, 2))
    gpu.set_input(O.VELOCITY_HDG_RATE_CMD, cmd[0::2], idx=np.arange(0, n, 2))
    orc.make_step(0.01, 800)
    for _ in range(800):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="landing")
    assert np.all(gpu.get_state()["x"][0::2, 2] == 0.0)


def test_feedforwards_are_sticky_and_ordered():
    n = 16
    orc, gpu = make_pair([af("x500")], None, grid_spawn(n, z=5.0))
    pos = _commands(O.POSITION_CMD, n)
    ffv = np.tile([0.5, -0.25, 0.1, 0.0], (n, 1))
    ffa = np.tile([0.2, 0.1, -0.1, 0.4], (n, 1))
    for s in (orc, gpu):
        s.set_input(O.POSITION_CMD, pos)
    # half get velocity_hdg_rate ff only, all get acceleration_hdg_rate ff, a quarter additionally acceleration_hdg
    orc.set_feedforward(3, ffv, None)
    gpu.set_feedforward("velocity_hdg_rate", ffv)
    orc.set_feedforward(0, ffa, None)
    gpu.set_feedforward("acceleration_hdg_rate", ffa)
    q = np.arange(0, n, 4)
    orc.set_feedforward(1, 2 * ffa[q], q)
    gpu.set_feedforward("acceleration_hdg", 2 * ffa[q], q)
    h = np.arange(0, n, 2)
    orc.set_feedforward(2, -ffv[h], h)
    gpu.set_feedforward("velocity_hdg", -ffv[h], h)
    orc.make_step(0.01, 300)
    for _ in range(300):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="ff position")
    vel = _commands(O.VELOCITY_HDG_RATE_CMD, n)
    for s in (orc, gpu):
        s.set_input(O.VELOCITY_HDG_RATE_CMD, vel)
    orc.make_step(0.01, 300)
    for _ in range(300):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="ff velocity_hdg_rate")


def test_crash_unknown_input_and_force():
    n = 8
    orc, gpu = make_pair([af("x500")], None, grid_spawn(n, z=50.0))
    cmd = _commands(O.VELOCITY_HDG_CMD, n)
    for s in (orc, gpu):
        s.set_input(O.VELOCITY_HDG_CMD, cmd)
        s.apply_force(np.tile([1.0, -2.0, 0.5], (n, 1)))
    orc.make_step(0.01, 100)
    for _ in range(100):
        gpu.make_step(0.01)
    for s in (orc, gpu):
        s.crash([1, 5])
        s.set_input(O.INPUT_UNKNOWN, None, [2])
        s.set_external_moment(np.tile([0.01, 0.0, -0.02], (2, 1)), [3, 4])
    orc.make_step(0.01, 100)
    for _ in range(100):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="crash")
    assert list(gpu.has_crashed()) == list(orc.has_crashed()) == [0, 1, 0, 0, 0, 1, 0, 0]
    rpm = gpu.get_state()["motor_rpm"]
    assert np.allclose(rpm[[1, 2, 5], :4], 1170.0, atol=50.0)  # idle RPM = rpm_min, not 0 (MM:408)


def test_set_params_resets_controllers_and_gains():
    n = 8
    orc, gpu = make_pair([af("x500")], None, grid_spawn(n, z=5.0))
    cmd = _commands(O.POSITION_CMD, n)
    for s in (orc, gpu):
        s.set_input(O.POSITION_CMD, cmd)
        s.set_controller_params("position", [1.5, 0.1, 0.1, 3.0], [0, 1, 2, 3])
        s.set_controller_params("velocity", [2.5, 0.04, 0.02, 3.0], [2, 3])
        s.set_controller_params("attitude", [5.0, 0.04, 0.02, 8.0, 0.8], [3, 4])
        s.set_controller_params("rate", [3.0, 0.03, 0.01], [4, 5])
        s.set_controller_params("mixer", [0.0], [5, 6])
    orc.make_step(0.01, 200)
    for _ in range(200):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="custom gains")
    heavy = af("x500", mass=2.6)
    for s in (orc, gpu):
        s.set_params(heavy, [0, 3, 7])  # uav_system.hpp:404-409: default gains again, PIDs reset
    assert gpu.get_controller_params(3).pos_kp == 2.0 and gpu.get_controller_params(2).pos_kp == 1.5
    assert gpu.get_params(0).mass == 2.6 and gpu.get_params(1).mass == 2.0
    orc.make_step(0.01, 300)
    for _ in range(300):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="after set_params")


def test_mixer_allocation_matches_reference_formula():
    for frame in ("x500", "f550", "naki", "t650"):
        orc, gpu = make_pair([af(frame)], None, np.zeros((1, 3)))
        assert np.max(np.abs(orc.get_mixer_allocation() - gpu.get_mixer_allocation())) < 1e-12, frame
    _, gpu = make_pair([af("x500")], None, np.zeros((1, 3)))
    m = gpu.get_mixer_allocation()[:4]
    s = np.sqrt(0.5)
    assert np.allclose(np.abs(m[:, :2]), s) and np.allclose(np.abs(m[:, 2]), 1.0) and np.allclose(m[:, 3], 1.0)


def test_uav_system_facade_single():
    from mrs_multirotor_simulator_b200 import UavSystem, model_params
    from mrs_multirotor_simulator_b200.uav_system import Position

    u = UavSystem(model_params(af("x500")), (0.0, 0.0, 1.0), 0.0)
    u.setInput(Position(position=np.array([5.0, -3.0, 4.0]), heading=1.0))
    for _ in range(2000):
        u.makeStep(0.005)
    st = u.getState()
    assert np.allclose(st.x, [4.99966878, -2.99316222, 3.99492282], atol=2e-8)
    assert abs(np.arctan2(st.R[1, 0], st.R[0, 0]) - 0.99999634) < 1e-7
    assert st.motor_rpm.shape == (4,) and not u.hasCrashed()


def test_pipelined_host_io_equals_blocking_io():
    """mrsb_set_input_async / mrsb_get_positions_async (upload and download on their own streams,
    double-buffered) deliver exactly what the blocking setInput / getState pair does."""
    import torch

    from mrs_multirotor_simulator_b200 import UavBatch

    n, ticks = 5000, 40
    spawn = grid_spawn(n, z=10.0)
    a = UavBatch([af("x500")], spawn_xyz=spawn, n=n)
    b = UavBatch([af("x500")], spawn_xyz=spawn, n=n)
    cmds = [torch.from_numpy(np.stack([rand(t, 1, n, -2, 2), rand(t, 2, n, -2, 2), rand(t, 3, n, -1, 1), rand(t, 4, n, -1, 1)], axis=1)).pin_memory()
            for t in range(ticks)]
    outs = [torch.zeros((n, 3), dtype=torch.float64).pin_memory() for _ in range(ticks)]
    ref = []
    for t in range(ticks):
        a.set_input(O.VELOCITY_HDG_RATE_CMD, cmds[t].numpy())
        a.make_step(0.01)
        ref.append(a.get_state(fields=("x",))["x"].copy())
    for t in range(ticks):  # nothing blocks here: all ticks are enqueued back to back
        b.set_input_async(O.VELOCITY_HDG_RATE_CMD, cmds[t].data_ptr(), 4)
        b.make_step(0.01)
        b.get_positions_async(outs[t].data_ptr())
    b.sync()
    for t in range(ticks):
        assert np.array_equal(outs[t].numpy(), ref[t]), t
    fa, fb = a.get_full_state(), b.get_full_state()
    for k in fa:
        assert np.array_equal(fa[k], fb[k]), k


def test_position_subset_download():
    """mrsb_set_position_subset: mrsb_get_positions_async then delivers the positions of the chosen UAVs only, in the chosen order
    (mixed airframes: the batch is bucketed internally, the subset is in the caller's indices); None switches back to all."""
    import torch

    from mrs_multirotor_simulator_b200 import UavBatch

    n = 3000
    tou = (np.arange(n) % 2).astype(np.int32)
    b = UavBatch([af("x500"), af("f550")], type_of_uav=tou, spawn_xyz=grid_spawn(n, z=5.0), n=n)
    b.set_input(O.VELOCITY_HDG_RATE_CMD, np.stack([rand(1, 1, n, -2, 2), rand(1, 2, n, -2, 2), rand(1, 3, n, -1, 1), rand(1, 4, n, -1, 1)], axis=1))
    sub = np.array([2999, 0, 17, 1500, 18, 1, 2998], dtype=np.int32)
    b.set_position_subset(sub)
    outs = [torch.zeros((len(sub), 3), dtype=torch.float64).pin_memory() for _ in range(6)]
    ref = []
    for t in range(6):
        b.make_step(0.01)
        b.get_positions_async(outs[t].data_ptr())
        ref.append(b.get_state(idx=sub, fields=("x",))["x"].copy())
    b.sync()
    for t in range(6):
        assert np.array_equal(outs[t].numpy(), ref[t]), t
    b.set_position_subset(None)
    full = torch.zeros((n, 3), dtype=torch.float64).pin_memory()
    b.make_step(0.01)
    b.get_positions_async(full.data_ptr())
    b.sync()
    assert np.array_equal(full.numpy(), b.get_state(fields=("x",))["x"])
