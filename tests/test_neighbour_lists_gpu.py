"""The collision pass keeps per-UAV neighbour lists between rebuilds of its spatial hash (collide.cu):
results must not depend on it.  Every check here compares against the reference's real nanoflann loop
(oracle/_ref) or against the same library with the lists switched off (MRSB_NO_NEIGHBOUR_LISTS=1: the
full pass every tick) — pair lists as exact sets, forces and trajectories bit for bit."""
import os

import numpy as np
import pytest

from helpers import grid_spawn, rand
from oracle import binding as O
from test_collisions_gpu import ENGINE, af, geometry, sorted_pairs

pytestmark = pytest.mark.gpu


def batch(types, tou, spawn, lists=True, **kw):
    from mrs_multirotor_simulator_b200 import UavBatch

    old = os.environ.pop("MRSB_NO_NEIGHBOUR_LISTS", None)
    if not lists:
        os.environ["MRSB_NO_NEIGHBOUR_LISTS"] = "1"
    try:
        b = UavBatch(types, type_of_uav=tou, spawn_xyz=spawn, n=len(spawn), **kw)
    finally:
        os.environ.pop("MRSB_NO_NEIGHBOUR_LISTS", None)
        if old is not None:
            os.environ["MRSB_NO_NEIGHBOUR_LISTS"] = old
    assert b.collision_info()["neighbour_lists"] == lists
    return b


def crowd(n, pitch, seed=11):
    side = int(np.ceil(np.sqrt(n)))
    k = np.arange(n)
    return np.stack([pitch * (k % side) + rand(seed, 0, n, -0.3, 0.3), pitch * (k // side) + rand(seed, 1, n, -0.3, 0.3), rand(seed, 2, n, 9.0, 10.5)], axis=1)


@pytest.mark.parametrize("crash", [False, True])
def test_lists_equal_full_pass_every_tick(crash):
    """2.2 m pitch crowd flying random velocities for 6 s: pairs every tick, forces, crash flags, end state."""
    n = 4096
    types = [af("x500"), af("f550")]
    tou = (np.arange(n) % 2).astype(np.int32)
    spawn = crowd(n, 2.2)
    cmd = np.stack([rand(5, 1, n, -3, 3), rand(5, 2, n, -3, 3), rand(5, 3, n, -0.5, 0.5), rand(5, 4, n, -1, 1)], axis=1)
    a, b = batch(types, tou, spawn, lists=True), batch(types, tou, spawn, lists=False)
    for s in (a, b):
        s.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
        s.set_collisions(True, crash, 100.0)
        s.set_pair_capacity(16 * n)
    total = 0
    for tick in range(600):
        for s in (a, b):
            s.make_step(0.01)
            s.handle_collisions()
        if tick % 7 == 0 or tick > 590:
            pa, pb = a.get_collision_pairs(), b.get_collision_pairs()
            assert np.array_equal(pa, pb), f"tick {tick}"
            assert np.array_equal(a.get_force(), b.get_force()), f"tick {tick}"
            total += len(pa)
    sa, sb = a.get_full_state(), b.get_full_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k], equal_nan=True), k
    assert np.array_equal(a.has_crashed(), b.has_crashed())
    info = a.collision_info()
    assert total > 100 and 0 < info["rebuilds"] < info["passes"], info  # the lists were really used between rebuilds


def test_lists_vs_nanoflann_with_teleports_fast_uavs_and_skipped_passes():
    """Everything that moves UAVs other than one stepping launch must invalidate the lists."""
    n = 900
    types = [af("x500")]
    tou = np.zeros(n, dtype=np.int32)
    spawn = grid_spawn(n, pitch=5.0, z=20.0)
    b = batch(types, tou, spawn)
    b.set_collisions(True, False, 100.0)
    b.set_input(O.VELOCITY_HDG_RATE_CMD, np.stack([rand(8, 1, n, -1, 1), rand(8, 2, n, -1, 1), np.zeros(n), np.zeros(n)], axis=1))
    arm, prop, mass = geometry(types, tou)

    def check(what):
        b.handle_collisions()
        x = b.get_state()["x"]
        ref_pairs, ref_forces, _ = O.collide_snapshot(x, arm, prop, mass, False, 100.0, engine=ENGINE)
        assert np.array_equal(sorted_pairs(ref_pairs), b.get_collision_pairs()), what
        assert np.allclose(ref_forces, b.get_force(), rtol=1e-12, atol=0), what
        return len(ref_pairs)

    for _ in range(30):
        b.make_step(0.01)
        assert check("calm") == 0
    # teleport 40 UAVs next to others (set_state between two passes, no stepping launch at all)
    st = b.get_state()
    idx = np.arange(0, 80, 2)
    x = st["x"][idx + 1] + np.array([0.4, 0.1, 0.05])
    b.set_state(idx=idx, x=x)
    assert check("teleport") >= 80
    b.make_step(0.01)
    assert check("after teleport") >= 40
    # several stepping launches between two passes
    for _ in range(5):
        b.make_step(0.01)
    check("5 launches, one pass")
    # K fused substeps per launch
    for _ in range(10):
        b.make_step(0.01, 10)
        check("K=10")
    # very fast UAVs: 60 m/s sideways -> 0.6 m per tick, far more than the skin allows for long
    st = b.get_state()
    v = st["v"].copy()
    v[::3, 0] += 60.0
    b.set_state(v=v)
    hits = 0
    for _ in range(60):
        b.make_step(0.01)
        hits += check("fast")
    assert hits > 0
    info = b.collision_info()
    assert info["rebuilds"] >= 30 and info["crowded_uavs"] == 0, info


def test_overcrowded_uavs_walk_the_table_and_recover():
    """More candidates than a list holds: those UAVs walk the stencil of the (ageing) table instead, on
    every pass until the next rebuild finds them less crowded — same results throughout."""
    n = 200
    types = [af("x500")]
    tou = np.zeros(n, dtype=np.int32)
    xyz = np.stack([rand(5, 0, n, -1.0, 1.0), rand(5, 1, n, -1.0, 1.0), rand(5, 2, n, 9.0, 11.0)], axis=1)
    b = batch(types, tou, xyz)
    b.set_pair_capacity(n * n)
    b.set_collisions(True, False, 100.0)
    arm, prop, mass = geometry(types, tou)
    b.handle_collisions()
    ref_pairs, _, _ = O.collide_snapshot(xyz, arm, prop, mass, False, 100.0, engine=ENGINE, cap=n * n)
    port_pairs, port_forces, _ = O.collide_snapshot(xyz, arm, prop, mass, False, 100.0, engine="port", cap=n * n)
    assert np.array_equal(sorted_pairs(ref_pairs), b.get_collision_pairs()) and len(ref_pairs) > 10 * n
    assert np.array_equal(port_forces, b.get_force())
    # blow the crowd apart and let it settle
    out = xyz - xyz.mean(axis=0)
    out[:, 2] = 0.0
    out = 6.0 * out / np.maximum(np.linalg.norm(out, axis=1, keepdims=True), 1e-3)
    b.set_input(O.VELOCITY_HDG_RATE_CMD, np.concatenate([out, np.zeros((n, 1))], axis=1))
    assert b.collision_info()["crowded_uavs"] == n
    seen_partial = False
    for tick in range(900):
        b.make_step(0.01)
        b.handle_collisions()
        if tick % 20 == 0:  # while the crowd dissolves: crowded and listed UAVs side by side, lists ageing between rebuilds
            x = b.get_state()["x"]
            ref_pairs, _, _ = O.collide_snapshot(x, arm, prop, mass, False, 100.0, engine=ENGINE, cap=n * n)
            assert np.array_equal(sorted_pairs(ref_pairs), b.get_collision_pairs()), f"tick {tick}"
            seen_partial |= 0 < b.collision_info()["crowded_uavs"] < n
    info = b.collision_info()
    assert seen_partial and info["crowded_uavs"] == 0 and info["rebuilds"] < info["passes"], info  # recovered
    x = b.get_state()["x"]
    ref_pairs, ref_forces, _ = O.collide_snapshot(x, arm, prop, mass, False, 100.0, engine=ENGINE, cap=n * n)
    assert np.array_equal(sorted_pairs(ref_pairs), b.get_collision_pairs())


def test_applied_forces_are_replaced_by_the_next_pass():
    """SIM:356-358: every pass writes every UAV's external force, also on list-only ticks."""
    n = 64
    b = batch([af("x500")], np.zeros(n, dtype=np.int32), grid_spawn(n, pitch=8.0, z=5.0))
    b.set_collisions(True, False, 100.0)
    for _ in range(3):
        b.make_step(0.01)
        b.handle_collisions()
    b.apply_force(np.ones((n, 3)))
    b.make_step(0.01)
    b.handle_collisions()
    assert not b.get_force().any()


def test_forces_written_through_the_device_view_are_replaced_after_forces_written():
    """ext_force written behind the library's back (RL loop through mrsb_get_device_view): after
    mrsb_forces_written the next pass replaces every force; a colliding pair keeps getting its force."""
    n = 256
    spawn = grid_spawn(n, pitch=8.0, z=5.0)
    spawn[1] = spawn[0] + np.array([0.3, 0.0, 0.0])  # one colliding pair
    b = batch([af("x500")], np.zeros(n, dtype=np.int32), spawn)
    b.set_collisions(True, False, 100.0)
    for _ in range(3):
        b.make_step(0.01)
        b.handle_collisions()
    f0 = b.get_force()
    assert f0[0].any() and f0[1].any() and not f0[2:].any()
    # overwrite every force on the device, tell the library, step once
    from cuda.bindings import runtime as cudart

    v = b.device_view()
    ones = np.ones(((n + v.tile - 1) // v.tile) * 3 * v.tile)
    b.sync()
    (err,) = cudart.cudaMemcpy(v.ext_force, ones.ctypes.data, ones.nbytes, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice)
    assert int(err) == 0
    assert np.all(b.get_force() == 1.0)
    b.forces_written()
    b.make_step(0.01)
    b.handle_collisions()
    f1 = b.get_force()
    assert f1[0].any() and f1[1].any() and not f1[2:].any()
    assert not np.any(f1 == 1.0)
