"""The rows either side of the path: the ROS wrapper's arithmetic — input-timeout hover
synthesis, odometry / IMU / rangefinder packing, set_mass / set_ground_z — GPU vs the CPU oracle."""
import numpy as np
import pytest

from helpers import assert_parity, grid_spawn, make_pair, rand
from oracle import binding as O

pytestmark = pytest.mark.gpu


def af(name, **kw):
    from mrs_multirotor_simulator_b200 import airframe

    return airframe(name, **kw)


def fly(orc, gpu, steps):
    orc.make_step(0.01, steps)
    for _ in range(steps // 10):
        gpu.make_step(0.01, 10)


def test_observation_packing():
    """publishOdometry / publishIMU / publishRangefinder (uav_system_ros.cpp:340-420)."""
    import torch

    n = 256
    types = [af("x500", ground_enabled=True, ground_z=-1.5), af("naki", ground_enabled=True, ground_z=0.5)]
    tou = (np.arange(n) % 2).astype(np.int32)
    orc, gpu = make_pair(types, tou, grid_spawn(n, z=6.0), rand(1, 0, n, -3, 3))
    cmd = np.stack([rand(1, 1, n, -3, 3), rand(1, 2, n, -3, 3), rand(1, 3, n, -1, 1), rand(1, 4, n, -1, 1)], axis=1)
    for s in (orc, gpu):
        s.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
    fly(orc, gpu, 150)
    # a few UAVs upside down / at odd attitudes to hit every quaternion branch and the inverted range
    R = orc.get_state()["R"].copy()
    R[0] = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=float).T.reshape(9)
    R[1] = np.array([[-1, 0, 0], [0, 1, 0], [0, 0, -1]], dtype=float).T.reshape(9)
    R[2] = np.array([[-1, 0, 0], [0, -1, 0], [0, 0, 1]], dtype=float).T.reshape(9)
    R[3] = np.array([[0, 0, 1], [0, 1, 0], [-1, 0, 0]], dtype=float).T.reshape(9)
    for s in (orc, gpu):
        s.set_state(idx=[0, 1, 2, 3], R=R[:4])
    od_o, od_g = orc.get_odometry(), gpu.get_odometry()
    assert np.max(np.abs(od_o - od_g)) <= 1e-9
    assert np.allclose(np.linalg.norm(od_g[:, 3:7], axis=1), 1.0, atol=1e-9)
    im_o, im_g = orc.get_imu(), gpu.get_imu()
    assert np.max(np.abs(im_o - im_g)) <= 1e-6
    rg_o, rg_g = orc.get_rangefinder(), gpu.get_rangefinder()
    assert np.max(np.abs(rg_o - rg_g)) <= 1e-9
    assert rg_g[0, 0] == 41.0 and rg_g[1, 0] == 41.0 and 0 < rg_g[5, 0] < 40.0
    # device-resident packing: odometry 13 | IMU acceleration 3 | range 1
    buf = torch.zeros((n, 17), dtype=torch.float64, device="cuda")
    gpu.pack_observations_device(buf.data_ptr(), 17)
    gpu.sync()
    got = buf.cpu().numpy()
    assert np.array_equal(got[:, :13], od_g) and np.array_equal(got[:, 13:16], im_g[:, 3:6]) and np.array_equal(got[:, 16:], rg_g)


def test_input_timeout_hover_synthesis_every_mode():
    """UavSystemRos::timeoutInput (uav_system_ros.cpp:474-647)."""
    from test_step_parity import ALL_MODES, _commands

    n = 64
    for mode in ALL_MODES + [O.INPUT_UNKNOWN]:
        orc, gpu = make_pair([af("x500")], None, grid_spawn(n, z=20.0), rand(2, 0, n, -3, 3))
        if mode != O.INPUT_UNKNOWN:
            cmd = _commands(mode, n)
            for s in (orc, gpu):
                s.set_input(mode, cmd)
        fly(orc, gpu, 100)
        half = np.arange(0, n, 2)
        orc.timeout_input(half)
        gpu.timeout_input(half)
        assert np.array_equal(gpu.get_input_mode(), np.full(n, mode))
        fly(orc, gpu, 200)
        if mode in (O.ACTUATOR_CMD, O.CONTROL_GROUP_CMD, O.ATTITUDE_RATE_CMD, O.ATTITUDE_CMD, O.TILT_HDG_RATE_CMD):
            so, sg = orc.get_state(), gpu.get_full_state()  # zero throttle: free fall, compare relative to the excursion
            for k in ("x", "v"):
                assert np.max(np.abs(so[k] - sg[k])) <= 1e-7 * (1.0 + np.max(np.abs(so[k]))), (mode, k)
        else:
            assert_parity(orc, gpu, what=f"timeout mode {mode}")
        if mode in (O.VELOCITY_HDG_CMD, O.VELOCITY_HDG_RATE_CMD):
            assert np.max(np.abs(gpu.get_state(half)["v"])) < 0.2, mode  # zero-velocity command: they came to a hover
            assert np.max(np.abs(gpu.get_state(half + 1)["v"])) > 0.5  # the others fly on


def test_set_mass_and_set_ground_z_services():
    """callbackSetMass / callbackSetGroundZ (uav_system_ros.cpp:1028-1080) go through setParams."""
    n = 32
    orc, gpu = make_pair([af("f550", ground_enabled=True, ground_z=0.0)], None, grid_spawn(n, z=4.0))
    cmd = np.tile([0.5, -0.5, 0.2, 0.3], (n, 1))
    for s in (orc, gpu):
        s.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
        s.set_controller_params("velocity", [1.5, 0.04, 0.02, 3.0])
    fly(orc, gpu, 100)
    heavy = np.arange(0, n, 3)
    mass = 2.3 + 0.05 * np.arange(len(heavy))
    low = np.arange(1, n, 3)
    for s in (orc, gpu):
        s.set_mass(mass, heavy)
        s.set_ground_z(np.full(len(low), 2.0), low)
    po, pg = orc.get_params(int(heavy[1])), gpu.get_params(int(heavy[1]))
    assert pg.mass == po.mass == mass[1] and list(pg.J) == list(po.J) and list(pg.allocation_matrix) == list(po.allocation_matrix)
    assert gpu.get_controller_params(int(heavy[1])).vel_kp == 2.0 and gpu.get_controller_params(2).vel_kp == 1.5  # setParams resets gains
    assert gpu.get_params(int(low[0])).ground_z == 2.0
    fly(orc, gpu, 300)
    assert_parity(orc, gpu, what="after set_mass / set_ground_z")
    for s in (orc, gpu):
        s.set_input(O.VELOCITY_HDG_RATE_CMD, np.tile([0.0, 0.0, -2.0, 0.0], (n, 1)))
    fly(orc, gpu, 600)
    assert_parity(orc, gpu, what="landed")
    z = gpu.get_state()["x"][:, 2]
    assert np.all(z[low] == 2.0) and np.all(z[heavy] == 0.0)


def test_device_resident_commands_and_state_view():
    """mrsb_set_input_device (commands produced on the GPU) and mrsb_get_device_view (tiled state arrays)."""
    import torch

    from mrs_multirotor_simulator_b200 import UavBatch

    n = 1000  # not a multiple of the 128-UAV tile
    spawn = grid_spawn(n, z=8.0)
    cmd = np.stack([rand(3, 1, n, -2, 2), rand(3, 2, n, -2, 2), rand(3, 3, n, -1, 1), rand(3, 4, n, -3, 3)], axis=1)
    a = UavBatch([af("x500")], spawn_xyz=spawn, n=n)
    b = UavBatch([af("x500")], spawn_xyz=spawn, n=n)
    a.set_input(O.VELOCITY_HDG_CMD, cmd)
    stream = torch.cuda.ExternalStream(b.stream)
    with torch.cuda.stream(stream):
        dev_cmd = torch.from_numpy(cmd).cuda()
    stream.synchronize()
    b.set_input_device(O.VELOCITY_HDG_CMD, dev_cmd.data_ptr(), 4)
    # a subset, addressed by a device index list, gets a different command on both
    sub = np.arange(5, n, 7).astype(np.int32)
    sub_cmd = np.tile([0.0, 0.0, 1.0, 0.5], (len(sub), 1))
    a.set_input(O.VELOCITY_HDG_RATE_CMD, sub_cmd, idx=sub)
    with torch.cuda.stream(stream):
        d_idx, d_sub = torch.from_numpy(sub).cuda(), torch.from_numpy(sub_cmd).cuda()
    stream.synchronize()
    b.set_input_device(O.VELOCITY_HDG_RATE_CMD, d_sub.data_ptr(), 4, n=len(sub), idx_ptr=d_idx.data_ptr())
    for s in (a, b):
        for _ in range(20):
            s.make_step(0.01, 5)
    sa, sb = a.get_full_state(), b.get_full_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    assert np.array_equal(a.get_input_mode(), b.get_input_mode()) and set(np.unique(b.get_input_mode())) == {8, 9}

    # tiled device view: component c of UAV i at ptr[((i // tile) * rows + c) * tile + i % tile]
    from cuda.bindings import runtime as cudart

    v = b.device_view()
    assert v.tile == 128 and v.state_rows == 18
    n_tiles = (n + v.tile - 1) // v.tile
    raw = np.zeros(n_tiles * v.state_rows * v.tile)
    b.sync()
    (err,) = cudart.cudaMemcpy(raw.ctypes.data, v.state, raw.nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    assert int(err) == 0
    tiles = raw.reshape(n_tiles, v.state_rows, v.tile)
    i = np.arange(n)
    for c in range(3):
        assert np.array_equal(tiles[i // v.tile, c, i % v.tile], sb["x"][:, c])
        assert np.array_equal(tiles[i // v.tile, 15 + c, i % v.tile], sb["omega"][:, c])
    for c in range(9):
        assert np.array_equal(tiles[i // v.tile, 6 + c, i % v.tile], sb["R"][:, c])


def test_tracker_cmd_sets_the_four_feedforwards():
    """UavSystemRos::callbackTrackerCmd (uav_system_ros.cpp:987-1022): mrsb_set_tracker_cmd == the four
    setFeedforward calls the callback makes, with the unused parts zeroed — against the reference's own code."""
    from mrs_multirotor_simulator_b200 import UavBatch

    n = 64
    spawn = grid_spawn(n, z=6.0)
    rows = np.stack([rand(4, 1, n, -1, 1), rand(4, 2, n, -1, 1), rand(4, 3, n, -0.5, 0.5), rand(4, 4, n, -1, 1), rand(4, 5, n, -1, 1),
                     rand(4, 6, n, -0.5, 0.5), rand(4, 7, n, -0.5, 0.5), (np.arange(n) % 2).astype(float), (np.arange(n) // 2 % 2).astype(float),
                     (np.arange(n) // 4 % 2).astype(float), (np.arange(n) // 8 % 2).astype(float)], axis=1)
    uh, uv, ur, ua = (rows[:, 7 + k] != 0 for k in range(4))
    v = np.stack([np.where(uh, rows[:, 0], 0.0), np.where(uh, rows[:, 1], 0.0), np.where(uv, rows[:, 2], 0.0)], axis=1)
    a = np.where(ua[:, None], rows[:, 3:6], 0.0)
    rate = np.where(ur, rows[:, 6], 0.0)
    zero = np.zeros(n)
    gpu = UavBatch([af("x500")], spawn_xyz=spawn, n=n)
    gpu.set_tracker_cmd(rows)
    have_ref = O.refsys_lib() is not None
    ref = (O.RefSwarm if have_ref else O.OracleSwarm)([af("x500")], spawn_xyz=spawn, n=n)
    ref.set_feedforward(2, np.column_stack([v, zero]))   # VelocityHdg(velocity, 0)
    ref.set_feedforward(3, np.column_stack([v, rate]))   # VelocityHdgRate(velocity, heading_rate)
    ref.set_feedforward(1, np.column_stack([a, zero]))   # AccelerationHdg(acceleration, 0)
    ref.set_feedforward(0, np.column_stack([a, rate]))   # AccelerationHdgRate(acceleration, heading_rate)
    for mode in (O.POSITION_CMD, O.VELOCITY_HDG_RATE_CMD, O.VELOCITY_HDG_CMD):
        cmd = np.stack([spawn[:, 0] + 1.0, spawn[:, 1] - 1.0, np.full(n, 7.0), rand(4, 9, n, -1, 1)], axis=1) if mode == O.POSITION_CMD else \
            np.stack([rand(4, 10, n, -1, 1), rand(4, 11, n, -1, 1), rand(4, 12, n, -0.5, 0.5), rand(4, 13, n, -1, 1)], axis=1)
        for s in (ref, gpu):
            s.set_input(mode, cmd)
        ref.make_step(0.01, 150)
        for _ in range(150):
            gpu.make_step(0.01)
        assert_parity(ref, gpu, what=f"tracker feed-forwards, mode {mode}")
