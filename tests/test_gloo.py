"""world_size-2 check of the host side of sharded operation on CPU (gloo): shard ranges, the
rendezvous that ships the communicator id from rank 0, and the packed-xyz gather layout the NCCL
all-gather fills on the GPUs (api.cu: exchange_positions)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mrs_multirotor_simulator_b200.sharding import connect, gather_layout, shard_range


class FakeBatch:
    """Stands in for UavBatch: records what connect() hands to mrsb_comm_init_nccl."""
    joined = None

    @staticmethod
    def nccl_unique_id():
        return bytes(range(128))

    def comm_init_nccl(self, world, rank, uid):
        self.joined = (world, rank, uid)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = FakeBatch()
    connect(b, dist)
    begin, count = shard_range(n, world, rank)
    k = np.arange(begin, begin + count)
    mine = torch.from_numpy(np.stack([4.0 * (k % 32), 4.0 * (k // 32), 0.5 * k], axis=1).reshape(-1))  # this shard's packed xyz
    lay = gather_layout(n, world)
    assert lay[rank] == (3 * begin, 3 * count)
    parts = [torch.empty(c, dtype=torch.float64) for _, c in lay]
    dist.all_gather(parts, mine)
    buf = torch.cat(parts).numpy().reshape(n, 3)
    # bench.py's state checksum: the shards' checksums combine (wrapping sum over ranks) to the checksum of the whole swarm
    import bench

    whole = np.stack([4.0 * (np.arange(n) % 32), 4.0 * (np.arange(n) // 32), 0.5 * np.arange(n)], axis=1) - 1e9 / 7.0  # negative values: top bits set
    combined = bench.combine_checksums(bench.checksum(whole[begin:begin + count]), dist)
    assert combined == bench.checksum(whole), (combined, bench.checksum(whole))
    q.put((rank, b.joined[0], b.joined[1], b.joined[2] == bytes(range(128)), float(np.abs(buf[:, 2] - 0.5 * np.arange(n)).max())))
    dist.destroy_process_group()


def test_two_ranks_rendezvous_and_gather_layout():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 1000  # divisible by 2 -> equal shards (single ncclAllGather path)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert got == [(0, 2, 0, True, 0.0), (1, 2, 1, True, 0.0)]
