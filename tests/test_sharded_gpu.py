"""Sharded operation on ONE GPU: two handles own the two halves of a swarm (contiguous index
shards), the packed positions are exchanged by copying each shard's slice of the
gather buffer into the other handle (what the NCCL all-gather does across GPUs), and the result
must equal the unsharded run bit for bit: pair lists, forces, crash flags, trajectories."""
import numpy as np
import pytest

from helpers import grid_spawn, rand
from oracle import binding as O

pytestmark = pytest.mark.gpu


def af(name, **kw):
    from mrs_multirotor_simulator_b200 import airframe

    return airframe(name, **kw)


def exchange(shards):
    """all-gather of the packed xyz between handles living on the same device"""
    from cuda.bindings import runtime as cudart

    for a in shards:
        a.sync()
    for dst in shards:
        pd, _ = dst.gather_buffer()
        for src in shards:
            if src is dst:
                continue
            ps, _ = src.gather_buffer()
            off = 24 * src.shard_begin
            (err,) = cudart.cudaMemcpy(pd + off, ps + off, 24 * src.n, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
            assert int(err) == 0


def make_shards(types, tou, spawn, cuts):
    from mrs_multirotor_simulator_b200 import UavBatch

    n = len(spawn)
    out = []
    for b, e in zip(cuts[:-1], cuts[1:]):
        out.append(UavBatch(types, type_of_uav=tou, spawn_xyz=spawn[b:e], n=e - b, n_global=n, shard_begin=b))
    return out


@pytest.mark.parametrize("crash", [False, True])
@pytest.mark.parametrize("cuts", [(0, 2048, 4096), (0, 1000, 1001, 4096)])
def test_sharded_collision_pass_equals_unsharded(crash, cuts):
    from mrs_multirotor_simulator_b200 import UavBatch

    n = 4096
    types = [af("x500"), af("naki"), af("t650")]
    tou = (np.arange(n) * 5 % 3).astype(np.int32)
    # well mixed: shard membership is unrelated to position
    xyz = np.stack([rand(11, 0, n, 0, 60), rand(11, 1, n, 0, 60), rand(11, 2, n, 0, 6)], axis=1)
    whole = UavBatch(types, type_of_uav=tou, spawn_xyz=xyz, n=n)
    whole.set_pair_capacity(8 * n)
    whole.set_collisions(True, crash, 100.0)
    whole.handle_collisions()
    ref_pairs, ref_forces, ref_crashed = whole.get_collision_pairs(), whole.get_force(), whole.has_crashed()
    assert len(ref_pairs) > 200

    shards = make_shards(types, tou, xyz, cuts)
    for sh in shards:
        sh.set_pair_capacity(8 * n)
        sh.set_collisions(True, crash, 100.0)
        sh.publish_positions()
    exchange(shards)
    pairs, forces, crashed = [], [], []
    for sh in shards:
        sh.handle_collisions_gathered()
        pairs.append(sh.get_collision_pairs())
        forces.append(sh.get_force())
        crashed.append(sh.has_crashed())
    pairs = np.concatenate(pairs)
    pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
    assert np.array_equal(ref_pairs, pairs)
    assert np.array_equal(ref_forces, np.concatenate(forces))
    assert np.array_equal(ref_crashed, np.concatenate(crashed))


def test_sharded_trajectory_with_halo_filter_equals_unsharded():
    """Spatially coherent shards (rows of a grid): only a thin halo of remote UAVs is inserted into
    each shard's table; 300 ticks of flight with rebounce must match the unsharded run exactly."""
    from mrs_multirotor_simulator_b200 import UavBatch

    n = 1024
    t = af("f550", ground_enabled=True)
    spawn = grid_spawn(n, pitch=1.2, z=3.0)
    cmd = np.stack([rand(5, 1, n, -2, 2), rand(5, 2, n, -2, 2), rand(5, 3, n, -0.5, 0.5), rand(5, 4, n, -1, 1)], axis=1)
    whole = UavBatch([t], spawn_xyz=spawn, n=n)
    shards = make_shards([t], None, spawn, (0, 512, 1024))
    whole.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
    for sh in shards:
        sh.set_input(O.VELOCITY_HDG_RATE_CMD, cmd[sh.shard_begin:sh.shard_begin + sh.n])
    for b in [whole] + shards:
        b.set_collisions(True, False, 100.0)
        b.set_pair_capacity(8 * n)
    total = 0
    for tick in range(300):
        whole.make_step(0.01)
        whole.handle_collisions()
        for sh in shards:
            sh.make_step(0.01)
        exchange(shards)
        for sh in shards:
            sh.handle_collisions_gathered()
        if tick % 25 == 24:
            p = np.concatenate([sh.get_collision_pairs() for sh in shards])
            p = p[np.lexsort((p[:, 1], p[:, 0]))] if len(p) else p
            assert np.array_equal(whole.get_collision_pairs(), p), tick
            total += len(p)
    assert total > 0
    sw = whole.get_full_state()
    for sh in shards:
        ss = sh.get_full_state()
        for k in sw:
            assert np.array_equal(sw[k][sh.shard_begin:sh.shard_begin + sh.n], ss[k]), k


def test_sharded_handle_without_exchange_is_an_error():
    from mrs_multirotor_simulator_b200 import UavBatch
    from mrs_multirotor_simulator_b200._lib import MrsbError

    b = UavBatch([af("x500")], spawn_xyz=np.zeros((4, 3)), n=4, n_global=8, shard_begin=4)
    b.set_collisions(True, False, 100.0)
    with pytest.raises(MrsbError):
        b.handle_collisions()
