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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
