"""GPU spatial-hash collision pass vs the reference's nanoflann KD-tree pass
(MultirotorSimulator::handleCollisions, src/multirotor_simulator.cpp:295-359).

Pair lists are compared as exact sets of directed (i, j) on IDENTICAL FP64 position snapshots
(SURVEY §8c): the snapshot is written into both sides, then both run one pass.
"""
import numpy as np
import pytest

from helpers import assert_parity, grid_spawn, make_pair, rand
from oracle import binding as O

pytestmark = pytest.mark.gpu

ENGINE = "nanoflann" if O.ref_lib() is not None else "port"


def af(name, **kw):
    from mrs_multirotor_simulator_b200 import airframe

    return airframe(name, **kw)


def snapshot(n, seed=42):
    """SURVEY §8d snapshot set: x=4i+U(-2,2), y=4j+U(-2,2), z=U(2,4) on the sqrt(N) grid."""
    side = int(np.ceil(np.sqrt(n)))
    k = np.arange(n)
    return np.stack([4.0 * (k % side) + rand(seed, 0, n, -2, 2), 4.0 * (k // side) + rand(seed, 1, n, -2, 2), rand(seed, 2, n, 2, 4)], axis=1)


def geometry(types, tou):
    arm = np.array([t["arm_length"] for t in types])[tou]
    prop = np.array([t["prop_radius"] for t in types])[tou]
    mass = np.array([t["mass"] for t in types])[tou]
    return arm, prop, mass


def run_gpu_pass(types, tou, xyz, crash, rebounce, pair_cap=None):
    from mrs_multirotor_simulator_b200 import UavBatch

    n = len(xyz)
    b = UavBatch(types, type_of_uav=tou, spawn_xyz=xyz, n=n)
    if pair_cap:
        b.set_pair_capacity(pair_cap)
    b.set_collisions(True, crash, rebounce)
    b.handle_collisions()
    return b, b.get_collision_pairs(), b.get_force(), b.has_crashed()


def sorted_pairs(p):
    p = np.asarray(p).reshape(-1, 2)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


@pytest.mark.parametrize("n", [400, 65536, 1048576])
def test_pair_list_bit_exact_f550_snapshots(n):
    types = [af("f550")]
    tou = np.zeros(n, dtype=np.int32)
    xyz = snapshot(n)
    if n == 400:
        xyz[:, :2] *= 0.5  # 2 m pitch so that the small case has pairs too
    arm, prop, mass = geometry(types, tou)
    ref_pairs, ref_forces, _ = O.collide_snapshot(xyz, arm, prop, mass, False, 100.0, engine=ENGINE, n_threads=8)
    b, pairs, forces, crashed = run_gpu_pass(types, tou, xyz, False, 100.0)
    assert len(ref_pairs) > 0
    assert np.array_equal(sorted_pairs(ref_pairs), pairs)
    assert np.array_equal(ref_forces, forces)  # <= 2 neighbours each here: sums are order independent
    assert not crashed.any()


def test_crash_mode_and_mixed_types():
    n = 20000
    types = [af(f) for f in ("x500", "f550", "naki", "t650", "robofly")]
    tou = (np.arange(n) * 7 % 5).astype(np.int32)
    xyz = snapshot(n, seed=7)
    xyz[:, :2] *= 0.35  # denser: ~1.4 m pitch
    arm, prop, mass = geometry(types, tou)
    ref_pairs, _, ref_crashed = O.collide_snapshot(xyz, arm, prop, mass, True, 100.0, engine=ENGINE, n_threads=8)
    b, pairs, forces, crashed = run_gpu_pass(types, tou, xyz, True, 100.0)
    assert len(ref_pairs) > 1000
    assert np.array_equal(sorted_pairs(ref_pairs), pairs)
    assert np.array_equal(ref_crashed.astype(np.int32), crashed)
    assert not forces.any()  # SIM:315-319,356-358: crash mode applies zero forces


def test_dense_cluster_many_neighbours():
    """All UAVs inside one search ball (>= 3 neighbours each): pair set exact, forces equal to
    rounding (ascending-j summation vs KD-tree traversal order)."""
    n = 300
    types = [af("t650"), af("x500")]
    tou = (np.arange(n) % 2).astype(np.int32)
    xyz = np.stack([rand(5, 0, n, -0.6, 0.6), rand(5, 1, n, -0.6, 0.6), rand(5, 2, n, 9.4, 10.6)], axis=1)
    xyz[10] = xyz[11]  # coincident distinct UAVs collide with zero force (normalized(0) = 0)
    arm, prop, mass = geometry(types, tou)
    ref_pairs, ref_forces, _ = O.collide_snapshot(xyz, arm, prop, mass, False, 100.0, engine=ENGINE, cap=n * n)
    port_pairs, port_forces, _ = O.collide_snapshot(xyz, arm, prop, mass, False, 100.0, engine="port", cap=n * n)
    b, pairs, forces, _ = run_gpu_pass(types, tou, xyz, False, 100.0, pair_cap=n * n)
    assert len(ref_pairs) > 10 * n
    assert np.array_equal(sorted_pairs(ref_pairs), pairs)
    assert np.array_equal(port_forces, forces)  # same summation order as the port: bit exact
    assert np.max(np.abs(ref_forces - forces)) <= 1e-9 * np.max(np.abs(ref_forces))


def test_edge_cases_far_apart_negative_and_huge_coordinates():
    types = [af("x500")]
    xyz = np.array([[0.0, 0.0, 0.0], [0.3, 0.0, 0.0], [-0.3, 0.0, 0.0], [-1000.2, -2000.1, -3.0], [-1000.2, -2000.5, -3.0], [1e9, 1e9, 1e9],
                    [1e9 + 0.25, 1e9, 1e9], [1e15, -1e15, 5.0], [1.99999, 0.0, 0.0], [2.00001, 0.0, 0.0], [-1.0e-9, 0.0, 0.5], [1.0e-9, 0.0, 0.5]])
    n = len(xyz)
    tou = np.zeros(n, dtype=np.int32)
    arm, prop, mass = geometry(types, tou)
    ref_pairs, ref_forces, _ = O.collide_snapshot(xyz, arm, prop, mass, False, 100.0, engine=ENGINE)
    b, pairs, forces, _ = run_gpu_pass(types, tou, xyz, False, 100.0)
    assert np.array_equal(sorted_pairs(ref_pairs), pairs)
    assert len(pairs) >= 10
    assert np.allclose(ref_forces, forces, rtol=1e-12, atol=0)


def test_single_uav_and_disabled():
    types = [af("x500")]
    b, pairs, forces, crashed = run_gpu_pass(types, np.zeros(1, dtype=np.int32), np.array([[1.0, 2.0, 3.0]]), True, 100.0)
    assert len(pairs) == 0 and not crashed.any()
    from mrs_multirotor_simulator_b200 import UavBatch

    b = UavBatch(types, spawn_xyz=np.zeros((4, 3)), n=4)
    b.apply_force(np.ones((4, 3)))
    b.set_collisions(False, False, 100.0)
    b.handle_collisions()  # SIM:299-301: returns before touching the forces
    assert np.array_equal(b.get_force(), np.ones((4, 3)))
    b.set_collisions(True, False, 100.0)
    b.handle_collisions()  # coincident UAVs: pairs, zero force
    assert len(b.get_collision_pairs()) == 12 and not b.get_force().any()


def test_c2_400_uav_scenario_with_rebounce():
    """BASELINE config 2: 400 f550 on the 20x20 grid (4 m pitch, z=0), ground on, collisions on with
    crash:false / rebounce 100, two 0.01 s warm-up steps, seeded VelocityHdgRate commands, 10 s."""
    n = 400
    t = af("f550", ground_enabled=True, ground_z=0.0)
    spawn = grid_spawn(n, pitch=4.0, z=0.0)
    orc, gpu = make_pair([t], None, spawn)
    zero = np.zeros((n, 8))
    for s in (orc, gpu):
        s.set_input(O.ACTUATOR_CMD, zero)  # uav_system_ros.cpp:223-232
        s.make_step(0.01, 2) if s is orc else (s.make_step(0.01), s.make_step(0.01))
        s.set_collisions(True, False, 100.0)
    cmd = np.stack([rand(42, 1, n, -2, 2), rand(42, 2, n, -2, 2), rand(42, 3, n, 0, 2), rand(42, 4, n, -1, 1)], axis=1)
    orc.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
    gpu.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
    total_pairs = 0
    for tick in range(1000):
        orc.make_step(0.01)
        po = orc.handle_collisions(engine=ENGINE)
        gpu.make_step(0.01)
        gpu.handle_collisions()
        if len(po) or tick % 100 == 99:
            pg = gpu.get_collision_pairs()
            assert np.array_equal(sorted_pairs(po), pg), f"tick {tick}"
            total_pairs += len(pg)
    assert_parity(orc, gpu, what="C2")
    assert total_pairs > 0


def _row_hash(cy, cz):
    """collide.cu row_hash (kept in step with the kernel; used only to aim a test at the table wrap)."""
    m = 0xFFFFFFFF
    h = ((cy & m) * 0x9E3779B1 & m) ^ (((cz & m) * 0x85EBCA77 + 0x165667B1) & m)
    h ^= h >> 15
    h = h * 0x2C1B3C6D & m
    h ^= h >> 12
    return h


def test_stencil_row_across_the_end_of_the_bucket_table():
    """Two x-adjacent cells whose buckets are B-1 and 0: found through the mirror bucket.  Cell edge and
    table size are read back from the library (they depend on how the pass is organised)."""
    from mrs_multirotor_simulator_b200 import UavBatch

    probe = UavBatch([af("x500")], spawn_xyz=np.zeros((24, 3)), n=24)
    info = probe.collision_info()
    cell, n_buckets = info["cell"], info["n_buckets"]
    hits = []
    for cy in range(-40, 40):
        for cz in range(0, 4):
            cx = (n_buckets - 1 - _row_hash(cy, cz)) % n_buckets  # bucket(cx) == B-1, bucket(cx+1) == 0
            if cx < 200:
                hits.append((cx, cy, cz))
    assert len(hits) >= 4
    xyz = []
    for cx, cy, cz in hits[:8]:
        xb = cell * (cx + 1)  # boundary between cell cx and cx+1
        y, z = cell * cy + 0.5 * cell, cell * cz + 0.5 * cell
        xyz += [[xb - 0.2, y, z], [xb + 0.2, y, z], [xb + 0.45, y + 0.1, z]]
    xyz = np.array(xyz)
    n = len(xyz)
    assert n == 24
    types = [af("x500")]
    tou = np.zeros(n, dtype=np.int32)
    arm, prop, mass = geometry(types, tou)
    ref_pairs, ref_forces, _ = O.collide_snapshot(xyz, arm, prop, mass, False, 100.0, engine=ENGINE)
    b, pairs, forces, _ = run_gpu_pass(types, tou, xyz, False, 100.0)
    assert b.collision_info()["n_buckets"] == n_buckets
    assert len(ref_pairs) >= 4 * len(hits[:8])
    assert np.array_equal(sorted_pairs(ref_pairs), pairs)
    assert np.array_equal(ref_forces, forces)
    # the same through the ageing table: one stepping launch, then a list-only pass
    b.make_step(0.01)
    b.handle_collisions()
    x = b.get_state()["x"]
    ref_pairs, _, _ = O.collide_snapshot(x, arm, prop, mass, False, 100.0, engine=ENGINE)
    assert np.array_equal(sorted_pairs(ref_pairs), b.get_collision_pairs())
