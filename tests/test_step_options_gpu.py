"""Options around the stepping launch that must not change what is computed:

  * `iterate_without_input` (uav_system_ros.cpp:265): with False, a UAV is stepped only once a command has arrived
    (`time_last_input_ > 0`) and stops again after an input timeout (uav_system_ros.cpp:249-260);
  * `mrsb_set_outputs`: dropping the IMU / packed-position rows leaves every other bit alone;
  * `mrsb_run` (one CUDA graph per tick) == make_step + handle_collisions issued one by one;
  * the batched `set_mass` / `set_ground_z` (uav_system_ros.cpp:1028-1080) against the oracle, many distinct values;
  * a non-orthonormal rotation written by setState is orthonormalised in every RK stage like the reference does.
"""
import numpy as np
import pytest

from helpers import TOL, assert_parity, grid_spawn, make_pair, rand
from oracle import binding as O

pytestmark = pytest.mark.gpu


def af(name, **kw):
    from mrs_multirotor_simulator_b200 import airframe

    return airframe(name, **kw)


def vel_cmd(n, seed=9):
    return np.stack([rand(seed, 1, n, -2, 2), rand(seed, 2, n, -2, 2), rand(seed, 3, n, 0, 2), rand(seed, 4, n, -1, 1)], axis=1)


@pytest.mark.parametrize("n", [300, 58001])
def test_iterate_without_input_false_freezes_uncommanded_uavs(n):
    """Half of the swarm gets a VelocityHdgRate command, the other half none.  Reference behaviour (ROSW:265): the commanded
    half flies exactly as usual (oracle), the other half does not move, fall, or change in any bit; after a command it joins in;
    after an input timeout it freezes again.  n = 58,001 exercises the persistent staged kernel's masking."""
    from mrs_multirotor_simulator_b200 import UavBatch

    x500 = af("x500", ground_enabled=False)
    spawn = grid_spawn(n, z=5.0)
    gpu = UavBatch([x500], spawn_xyz=spawn, n=n)
    gpu.set_iterate_without_input(False)
    cmd_idx = np.arange(0, n, 2, dtype=np.int32)
    idle_idx = np.arange(1, n, 2, dtype=np.int32)
    cmd = vel_cmd(len(cmd_idx))
    before = gpu.get_full_state(idle_idx)
    gpu.set_input(O.VELOCITY_HDG_RATE_CMD, cmd, idx=cmd_idx)
    orc = O.OracleSwarm([x500], spawn_xyz=spawn[cmd_idx], n=len(cmd_idx))
    orc.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
    for _ in range(50):
        gpu.make_step(0.01)
    orc.make_step(0.01, 50, n_threads=8)
    after = gpu.get_full_state(idle_idx)
    for k in before:
        assert np.array_equal(before[k], after[k]), f"an uncommanded UAV changed its {k}"
    so, sg = orc.get_state(), gpu.get_full_state(cmd_idx)
    for k, t in TOL.items():
        assert np.max(np.abs(so[k] - sg[k])) <= t, k
    # the idle half receives a command: from now on it is stepped (free fall would show at once: it holds altitude instead)
    gpu.set_input(O.VELOCITY_HDG_RATE_CMD, np.zeros((len(idle_idx), 4)), idx=idle_idx)
    for _ in range(20):
        gpu.make_step(0.01)
    moved = gpu.get_full_state(idle_idx)
    assert np.all(moved["motor_rpm"][:, :4] > 0.0) and not np.array_equal(moved["v"], before["v"])
    # a fresh command for EVERYBODY in one call (uniform mode: at 58,001 UAVs the persistent staged kernel runs from here on), then
    # an input timeout on half of them (ROSW:249-260): hover command, time_last_input_ = 0 -> frozen again, the others fly on
    allcmd = vel_cmd(n, seed=21)
    gpu.set_input(O.VELOCITY_HDG_RATE_CMD, allcmd)
    gpu.timeout_input(idx=cmd_idx)
    frozen = gpu.get_full_state(cmd_idx)
    flying = gpu.get_state(idle_idx)
    for _ in range(10):
        gpu.make_step(0.01)
    if n > 57000:
        assert gpu.step_info()["variant"] == "staged", gpu.step_info()
    again = gpu.get_full_state(cmd_idx)
    for k in frozen:
        assert np.array_equal(frozen[k], again[k]), k
    assert not np.array_equal(flying["x"], gpu.get_state(idle_idx)["x"])
    # default (True): everybody is stepped, commanded or not — uncommanded UAVs drive their motors to zero and fall (US:308-310)
    gpu.set_iterate_without_input(True)
    gpu.make_step(0.01)
    assert not np.array_equal(gpu.get_state(cmd_idx)["x"], again["x"])


@pytest.mark.parametrize("n", [500, 58001])
def test_outputs_off_changes_nothing_else(n):
    """IMU rows and packed positions switched off: state, PIDs (through the next steps) and motor speeds are the same bits; the
    IMU getters refuse instead of returning stale numbers; switching positions back on publishes them for the collision pass."""
    from mrs_multirotor_simulator_b200 import UavBatch
    from mrs_multirotor_simulator_b200._lib import MrsbError

    x500 = af("x500", ground_enabled=True, ground_z=0.0)
    spawn = grid_spawn(n, pitch=1.5, z=3.0)
    cmd = vel_cmd(n)
    a = UavBatch([x500], spawn_xyz=spawn, n=n)
    b = UavBatch([x500], spawn_xyz=spawn, n=n)
    b.set_outputs(imu=False, positions=False)
    for s in (a, b):
        s.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
    for _ in range(60):
        a.make_step(0.01)
        b.make_step(0.01)
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    with pytest.raises(MrsbError):
        b.get_imu_acceleration()
    with pytest.raises(MrsbError):
        b.get_imu()
    # collisions need the positions: enabling them brings the packed buffer up to date, and the passes agree
    for s in (a, b):
        s.set_collisions(True, False, 100.0)
        s.set_pair_capacity(16 * n)
    for _ in range(30):
        for s in (a, b):
            s.make_step(0.01)
            s.handle_collisions()
    assert np.array_equal(a.get_collision_pairs(), b.get_collision_pairs())
    assert np.array_equal(a.get_force(), b.get_force())
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    b.set_outputs(imu=True, positions=True)
    a.make_step(0.01)
    b.make_step(0.01)
    assert np.array_equal(a.get_imu_acceleration(), b.get_imu_acceleration())


@pytest.mark.parametrize("crash", [False, True])
def test_run_tick_graph_equals_separate_calls(crash):
    """mrsb_run(n_ticks) — stepping launch and collision pass of every tick as one graph launch — against the same ticks issued as
    make_step + handle_collisions: pair lists, forces, crash flags and the full state, bit for bit, with rebuilds in between and a
    teleport that forces one."""
    from mrs_multirotor_simulator_b200 import UavBatch

    n = 6000
    f550 = af("f550", ground_enabled=True, ground_z=0.0)
    spawn = grid_spawn(n, pitch=2.0, z=4.0) + np.stack([rand(4, 0, n, -0.2, 0.2), rand(4, 1, n, -0.2, 0.2), rand(4, 2, n, -0.5, 0.5)], axis=1)
    cmd = vel_cmd(n, seed=12)
    a = UavBatch([f550], spawn_xyz=spawn, n=n)
    b = UavBatch([f550], spawn_xyz=spawn, n=n)
    for s in (a, b):
        s.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
        s.set_collisions(True, crash, 100.0)
        s.set_pair_capacity(16 * n)
    total = 0
    for block in range(6):
        for _ in range(25):
            a.make_step(0.01)
            a.handle_collisions()
        b.run(0.01, 25)
        pa, pb = a.get_collision_pairs(), b.get_collision_pairs()
        assert np.array_equal(pa, pb), block
        total += len(pa)
        if block == 2:
            idx = np.arange(0, n, 50, dtype=np.int32)
            for s in (a, b):
                s.set_state(idx=idx, x=spawn[idx] + np.array([0.7, 0.0, 0.3]))
    sa, sb = a.get_full_state(), b.get_full_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k], equal_nan=True), k
    assert np.array_equal(a.get_force(), b.get_force()) and np.array_equal(a.has_crashed(), b.has_crashed())
    assert total > 0
    ia, ib = a.collision_info(), b.collision_info()
    assert ia["rebuilds"] == ib["rebuilds"] and 0 < ib["rebuilds"] < ib["passes"], (ia, ib)
    assert b.counters()["steps"] == a.counters()["steps"] and b.counters()["collision_passes"] == a.counters()["collision_passes"]


def test_set_mass_and_ground_z_batched_many_values_vs_oracle():
    """set_mass / set_ground_z with a different value for every UAV (one call, one re-pointing), then 3 s of flight vs the oracle."""
    n = 700
    types = [af("x500", ground_enabled=True, ground_z=0.0), af("f550", ground_enabled=True, ground_z=0.0)]
    tou = (np.arange(n) % 2).astype(np.int32)
    orc, gpu = make_pair(types, tou, grid_spawn(n, z=2.0))
    cmd = vel_cmd(n, seed=3)
    mass = rand(6, 0, n, 1.5, 3.5)
    gz = rand(6, 1, n, -1.0, 1.5)
    for s in (orc, gpu):
        s.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
        s.make_step(0.01, 20) if s is orc else [s.make_step(0.01) for _ in range(20)]
        s.set_mass(mass)
        s.set_ground_z(gz[::3], idx=np.arange(0, n, 3, dtype=np.int32))
    orc.make_step(0.01, 300, n_threads=8)
    for _ in range(300):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="after batched set_mass / set_ground_z")
    for i in (0, 1, 17, n - 1):
        po, pg = orc.get_params(i), gpu.get_params(i)
        assert po.mass == pg.mass == mass[i] and tuple(po.J) == tuple(pg.J)
    # the same mass again for everybody: the unreferenced parameter sets are collected, results unchanged
    for s in (orc, gpu):
        s.set_mass(np.full(n, 2.5))
    orc.make_step(0.01, 50, n_threads=8)
    for _ in range(50):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="after uniform set_mass")


def test_raw_rotation_from_set_state_is_orthonormalised_in_every_stage():
    """A rotation written by setState may be anything (MM:424-433 stores it as is); the reference orthonormalises it in EVERY RK
    stage (MM:314-316), and so does the kernel."""
    n = 128
    orc, gpu = make_pair([af("x500")], None, grid_spawn(n, z=5.0))
    st = orc.get_state()
    R = st["R"] * (1.0 + np.stack([rand(2, c, n, -0.05, 0.05) for c in range(9)], axis=1))  # visibly non-orthonormal
    for s in (orc, gpu):
        s.set_state(R=R)
        s.set_input(O.VELOCITY_HDG_RATE_CMD, vel_cmd(n))
        s.make_step(0.01)
    tight = {"x": 1e-12, "v": 1e-11, "R": 1e-13, "omega": 1e-9, "motor_rpm": 1e-8}
    assert_parity(orc, gpu, tol=tight, what="first step after a raw R")
    orc.make_step(0.01, 300)
    for _ in range(300):
        gpu.make_step(0.01)
    assert_parity(orc, gpu, what="3 s after a raw R")


def _mixed_flight(bucketed, n, crash):
    """A mixed x500 / f550 / naki swarm (interleaved) through most of the API, with or without the bucketing by airframe."""
    import os

    from mrs_multirotor_simulator_b200 import UavBatch

    if bucketed:
        os.environ.pop("MRSB_NO_BUCKETS", None)
    else:
        os.environ["MRSB_NO_BUCKETS"] = "1"
    try:
        types = [af("x500", ground_enabled=True, ground_z=0.0), af("f550", ground_enabled=True, ground_z=0.0), af("naki", ground_enabled=True, ground_z=0.0)]
        tou = (np.arange(n) * 7 % 3).astype(np.int32)
        spawn = grid_spawn(n, pitch=1.6, z=3.0) + np.stack([rand(14, 0, n, -0.2, 0.2), rand(14, 1, n, -0.2, 0.2), rand(14, 2, n, -0.4, 0.4)], axis=1)
        b = UavBatch(types, type_of_uav=tou, spawn_xyz=spawn, spawn_heading=rand(14, 3, n, -3, 3), n=n)
        assert (b.device_view().slot_of_uav is not None) == bucketed
        b.set_collisions(True, crash, 100.0)
        b.set_pair_capacity(16 * n)
        b.set_input(O.VELOCITY_HDG_RATE_CMD, vel_cmd(n, seed=15))
        some = np.arange(3, n, 11, dtype=np.int32)
        b.set_input(O.POSITION_CMD, np.concatenate([spawn[some] + 2.0, rand(16, 0, len(some), -3, 3)[:, None]], axis=1), idx=some)
        b.set_feedforward("velocity_hdg", np.tile([0.1, -0.1, 0.0, 0.0], (len(some), 1)), idx=some)
        pairs = 0
        for t in range(120):
            b.make_step(0.01)
            b.handle_collisions()
            if t == 40:
                b.set_mass(2.0 + (some % 4) * 0.3, idx=some)
                b.apply_force(np.tile([0.5, 0.0, 0.2], (len(some), 1)), idx=some)
                b.set_state(idx=some[:50], x=spawn[some[:50]] + np.array([0.3, 0.3, 1.0]))
                b.crash(some[-5:])
            if t == 80:
                b.timeout_input(idx=some)
                b.set_controller_params("velocity", [2.5, 0.06, 0.02, 4.0], idx=some)
            pairs += len(b.get_collision_pairs())
        out = b.get_full_state()
        out["force"] = b.get_force()
        out["crashed"] = b.has_crashed()
        out["mode"] = b.get_input_mode()
        out["odom"] = b.get_odometry()
        out["range"] = b.get_rangefinder()
        out["pairs_last"] = b.get_collision_pairs()
        out["mass"] = np.array([b.get_params(int(i)).mass for i in some[:20]])
        b.close()
        return out, pairs
    finally:
        os.environ.pop("MRSB_NO_BUCKETS", None)


@pytest.mark.parametrize("crash", [False, True])
def test_bucketing_by_airframe_changes_no_bit(crash):
    """Batches with several airframes are stored sorted by airframe (whole tiles per airframe) so that the specialised kernels
    run; everything a caller can see — state, forces, crash flags, pair lists, observations, per-UAV parameters, addressed by the
    caller's own indices — must equal the unbucketed library (MRSB_NO_BUCKETS=1) bit for bit."""
    a, pa = _mixed_flight(True, 3000, crash)
    b, pb = _mixed_flight(False, 3000, crash)
    assert pa == pb and pa > 0
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
