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


def attach_engine(engine, transport=None):
    """Join `engine` to the initialised torch.distributed group.  transport: "p2p" (default: CUDA-IPC peer memory over NVLink;
    the L tail becomes ONE kernel that sums the per-CTA partials, the peers' contributions and the initial term) or "nccl";
    DLRA_COMM overrides.  If any rank cannot map its peers (no IPC / no peer access) every rank falls back to NCCL.
    A no-op for world size 1."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    transport = os.environ.get("DLRA_COMM", transport or "p2p")
    world, rank = dist.get_world_size(), dist.get_rank()
    if transport == "p2p":
        ok = True
        try:
            handle = engine.p2p_export()
        except Exception:
            handle, ok = b"\0" * 64, False
        handles = [None] * world
        dist.all_gather_object(handles, (handle, ok))
        if all(h[1] for h in handles):
            try:
                engine.p2p_import(world, rank, [h[0] for h in handles])
            except Exception:
                ok = False
        else:
            ok = False
        oks = [None] * world
        dist.all_gather_object(oks, ok)
        if all(oks):
            return "p2p"
        transport = "nccl"
    engine.comm_init(*comm_from_torch())
    return transport
