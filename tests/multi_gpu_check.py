#!/usr/bin/env python
"""Run under torchrun on >= 2 GPUs: the sharded swarm (fused peer-store exchange, then again with
the NCCL all-gather) must reproduce the unsharded run on one GPU bit for bit — pair lists every
tick a pair exists, forces, crash flags and the full state after 60 ticks.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import rand  # noqa: E402
from mrs_multirotor_simulator_b200 import VELOCITY_HDG_RATE_CMD, UavBatch, airframe  # noqa: E402
from mrs_multirotor_simulator_b200.sharding import connect, shard_range  # noqa: E402


VERBOSE = bool(os.environ.get("MGC_VERBOSE"))


def log(*a):
    if VERBOSE:
        print(f"[r{os.environ['RANK']}]", *a, flush=True)


def run(world, rank, local, crash, ticks=60, n=12001, area=90.0, every=10):
    """area 90 m: a crowd (most UAVs have more candidates than a list holds and walk the table instead); area 700 m: a sparse
    swarm whose neighbour lists live for many ticks and contain UAVs of other shards."""
    ticks = int(os.environ.get("MGC_TICKS", ticks))
    n = int(os.environ.get("MGC_N", n))
    log("run crash", crash, "no_p2p", os.environ.get("MRSB_NO_P2P"))
    types = [airframe("x500", ground_enabled=True), airframe("f550", ground_enabled=True), airframe("naki", ground_enabled=True)]
    tou = (np.arange(n) * 7 % 3).astype(np.int32)
    xyz = np.stack([rand(21, 0, n, 0, area), rand(21, 1, n, 0, area), rand(21, 2, n, 1, 7)], axis=1)  # well mixed: shards overlap everywhere
    cmd = np.stack([rand(22, 1, n, -2, 2), rand(22, 2, n, -2, 2), rand(22, 3, n, -0.5, 0.5), rand(22, 4, n, -1, 1)], axis=1)
    begin, count = shard_range(n, world, rank)
    mine = UavBatch(types, type_of_uav=tou, spawn_xyz=xyz[begin:begin + count], n=count, device=local, n_global=n, shard_begin=begin)
    connect(mine, dist)
    log("connected, mode", mine.exchange_mode())
    mine.set_input(VELOCITY_HDG_RATE_CMD, cmd[begin:begin + count])
    mine.set_collisions(True, crash, 100.0)
    mine.set_pair_capacity(4 * n)
    whole = None
    if rank == 0:
        whole = UavBatch(types, type_of_uav=tou, spawn_xyz=xyz, n=n, device=local)
        whole.set_input(VELOCITY_HDG_RATE_CMD, cmd)
        whole.set_collisions(True, crash, 100.0)
        whole.set_pair_capacity(4 * n)
    n_pairs = 0
    for t in range(ticks):
        mine.make_step(0.01)
        mine.handle_collisions()
        if rank == 0:
            whole.make_step(0.01)
            whole.handle_collisions()
        if t % every == every - 1:
            p = mine.get_collision_pairs()
            log("tick", t, "pairs", len(p))
            gathered = [None] * world
            dist.all_gather_object(gathered, p)
            if rank == 0:
                allp = np.concatenate(gathered)
                allp = allp[np.lexsort((allp[:, 1], allp[:, 0]))] if len(allp) else allp
                ref = whole.get_collision_pairs()
                assert np.array_equal(ref, allp), f"pair lists differ at tick {t}"
                n_pairs += len(ref)
    log("loop done")
    st = mine.get_full_state()
    st["crashed"] = mine.has_crashed()
    st["force"] = mine.get_force()
    gathered = [None] * world
    dist.all_gather_object(gathered, st)
    if rank == 0:
        ref = whole.get_full_state()
        ref["crashed"] = whole.has_crashed()
        ref["force"] = whole.get_force()
        for k in ref:
            got = np.concatenate([g[k] for g in gathered])
            assert np.array_equal(ref[k], got), f"{k} differs from the unsharded run"
        assert n_pairs > 0
    mode = mine.exchange_mode()
    info = mine.collision_info()
    assert info["neighbour_lists"] == (mode == 2), info  # lists need the displacement bound of every rank: carried by the peer hand-shake only
    if mode == 2 and area > 500:
        assert 0 < info["rebuilds"] < info["passes"] // 3 and info["crowded_uavs"] == 0, info  # the lists were really used between rebuilds
    log("run ok", info)
    dist.barrier()
    mine.close()
    if whole is not None:
        whole.close()
    dist.barrier()
    return mode, n_pairs


def checksum(a):
    """64-bit checksum of the raw bytes of an array (sum of its uint64 words, wrapping)."""
    return int(np.ascontiguousarray(a).view(np.uint64).sum(dtype=np.uint64))


def run_big(world, rank, local, per_rank=None, ticks=60):
    """The kernels bench.py --gpus N times: a uniform x500 swarm with more than 57k UAVs PER RANK (the persistent TMA-staged stepping
    kernel on every rank, asserted), neighbour lists across shards over the pull exchange — against the unsharded run on rank 0,
    bit for bit.  Then the things the exchange has to survive: two stepping launches between passes, a pass without a step, a
    teleport on one rank only, and a mass change (set_mass: the peers' collision pass reads the owner's geometry)."""
    per_rank = int(os.environ.get("MGC_PER_RANK", per_rank or 58001))
    n = per_rank * world
    x500 = airframe("x500", ground_enabled=True, ground_z=0.0, takeoff_patch_enabled=False)
    side = int(np.ceil(np.sqrt(n)))
    k = np.arange(n)
    # rows of a 1.9 m grid with jitter: shards are bands of rows (thin halos), neighbours closer than sqrt(3) exist from the start
    xyz = np.stack([1.9 * (k % side) + rand(31, 0, n, -0.25, 0.25), 1.9 * (k // side) + rand(31, 1, n, -0.25, 0.25), rand(31, 2, n, 2.0, 2.6)], axis=1)
    cmd = np.stack([rand(32, 1, n, -2, 2), rand(32, 2, n, -2, 2), rand(32, 3, n, -0.3, 0.3), rand(32, 4, n, -1, 1)], axis=1)
    begin, count = shard_range(n, world, rank)
    sl = slice(begin, begin + count)
    mine = UavBatch([x500], spawn_xyz=xyz[sl], n=count, device=local, n_global=n, shard_begin=begin)
    connect(mine, dist)
    assert mine.exchange_mode() == 2, "needs peer access"
    whole = UavBatch([x500], spawn_xyz=xyz, n=n, device=local) if rank == 0 else None
    for b, c in ((mine, cmd[sl]), (whole, cmd)):
        if b is not None:
            b.set_input(VELOCITY_HDG_RATE_CMD, c)
            b.set_collisions(True, False, 100.0)
            b.set_pair_capacity(8 * n)
    both = [b for b in (mine, whole) if b is not None]

    def compare(tag):
        st = mine.get_full_state()
        st["force"] = mine.get_force()
        st["crashed"] = mine.has_crashed()
        st["pairs"] = mine.get_collision_pairs()
        gathered = [None] * world
        dist.all_gather_object(gathered, st)
        if rank == 0:
            ref = whole.get_full_state()
            ref["force"] = whole.get_force()
            ref["crashed"] = whole.has_crashed()
            for key in ref:
                got = np.concatenate([g[key] for g in gathered])
                assert np.array_equal(ref[key], got, equal_nan=True), f"{tag}: {key} differs from the unsharded run"
            pairs = np.concatenate([g["pairs"] for g in gathered])
            pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))] if len(pairs) else pairs
            assert np.array_equal(whole.get_collision_pairs(), pairs), f"{tag}: pair lists differ"
            return len(pairs), checksum(ref["x"])
        return 0, 0

    total = 0
    for t in range(ticks):
        for b in both:
            b.make_step(0.01)
            b.handle_collisions()
        if t % 20 == 19:
            total += compare(f"tick {t}")[0]
    info = mine.step_info()
    assert info["variant"] == "staged" and info["mode"] == VELOCITY_HDG_RATE_CMD, info
    ci = mine.collision_info()
    assert ci["neighbour_lists"] and 0 < ci["rebuilds"] < ci["passes"], ci
    # mrsb_run: the whole tick as one graph launch per tick
    for b in both:
        b.run(0.01, 15)
    total += compare("after mrsb_run")[0]
    # two stepping launches between passes, then a pass without any step
    for b in both:
        b.make_step(0.01)
        b.make_step(0.01)
        b.handle_collisions()
        b.handle_collisions()
    total += compare("two steps per pass")[0]
    # a teleport on the last rank only (its peers must see the new positions at the next pass), without a step in between
    far = np.arange(n - 64, n)
    newpos = xyz[far] + np.array([0.4, -0.3, 0.2])
    if rank == world - 1:
        mine.set_state(idx=(far - begin).astype(np.int32), x=newpos)
    if whole is not None:
        whole.set_state(idx=far.astype(np.int32), x=newpos)
    for b in both:
        b.handle_collisions()
    total += compare("teleport")[0]
    # heavier UAVs along every shard boundary: the rebounce weight m_j / (m_i + m_j) of REMOTE neighbours changes (SIM:350)
    heavy = np.arange(0, n, 7)
    mass = 2.0 + (heavy % 5) * 0.5
    loc = heavy[(heavy >= begin) & (heavy < begin + count)]
    mine.set_mass(2.0 + (loc % 5) * 0.5, idx=(loc - begin).astype(np.int32))
    if whole is not None:
        whole.set_mass(mass, idx=heavy.astype(np.int32))
    for t in range(30):
        for b in both:
            b.make_step(0.01)
            b.handle_collisions()
    n_pairs, cs = compare("after set_mass")
    total += n_pairs
    assert rank != 0 or total > 0
    dist.barrier()
    mine.close()
    if whole is not None:
        whole.close()
    dist.barrier()
    return n, total, cs


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ.pop("MRSB_NO_P2P", None)
    if os.environ.get("MGC_SKIP_BIG") is None:
        n, pairs, cs = run_big(world, rank, local)
        if rank == 0:
            print("MULTI_GPU_BIG_OK world=%d n=%d staged kernel on every rank, pairs=%d, checksum(x)=%016x" % (world, n, pairs, cs), flush=True)
    results = []
    for no_p2p in ("", "1"):
        if no_p2p:
            os.environ["MRSB_NO_P2P"] = "1"
        else:
            os.environ.pop("MRSB_NO_P2P", None)
        for crash in (False, True):
            results.append(run(world, rank, local, crash))
        results.append(run(world, rank, local, False, ticks=400, area=700.0, every=25))
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK world=%d exchange_modes=%s pairs=%s" % (world, [m for m, _ in results], [p for _, p in results]), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
