"""Row sharding helpers: one process per GPU, contiguous row blocks of ΔA / U / K, NCCL inside libdlra.so.
torch.distributed is used only as plumbing to hand the ncclUniqueId to every rank."""
import os


def row_shard(n, world, rank):
    """Contiguous row block [lo, hi) of rank `rank`; blocks differ by at most one row and are even-sized when n/world is
    even (the TMA fast path wants an even local row count)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def comm_from_torch():
    """(nranks, rank, unique_id) for `init(..., comm=...)`, broadcasting the id made on rank 0 through torch.distributed."""
    import torch.distributed as dist
    from .engine import Engine
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    box = [Engine.nccl_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return dist.get_world_size(), dist.get_rank(), box[0]


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
