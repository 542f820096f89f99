"""Host-side helpers for sharded (one handle per GPU) operation: SURVEY §8e.

UAVs are partitioned into contiguous global index ranges [begin, begin+count), one per rank; equal
shards (n divisible by world) let the library use a single in-place ncclAllGather per tick, uneven
ones fall back to a group of broadcasts (api.cu: exchange_positions)."""


def shard_range(n_global, world_size, rank):
    """Contiguous shard of `rank`: the first (n_global % world_size) ranks get one extra UAV."""
    if not 0 <= rank < world_size:
        raise ValueError("rank outside 0..world_size-1")
    base, extra = divmod(int(n_global), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def gather_layout(n_global, world_size):
    """(offset, length) in doubles of every rank's slice of the packed-xyz gather buffer."""
    return [(3 * b, 3 * c) for b, c in (shard_range(n_global, world_size, r) for r in range(world_size))]


def connect(batch, dist, device=None):
    """Create the in-library NCCL communicator of `batch` from an initialised torch.distributed
    process group: rank 0 makes the unique id, everybody receives it, everybody joins."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    uid = [type(batch).nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0, device=device)
    batch.comm_init_nccl(world, rank, uid[0])
