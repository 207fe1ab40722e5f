"""Multi-GPU plumbing: one process per GPU (torchrun), source points sharded by rank, target replicated.

torch.distributed is used only to agree on the NCCL unique id (works over any backend, e.g. gloo on CPU); the
per-evaluation all-reduce of the 14 partial sums is issued by the engine itself on its own NCCL communicator
(csrc/engine.cu run_cost), so no Python sits on the evaluation path.
"""
import os


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of `n` Morton-sorted source points owned by `rank` (csrc/engine.cu update_shard)."""
    return n * rank // world, n * (rank + 1) // world


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def exchange_unique_id(make_id, rank, group=None):
    """Rank 0 calls make_id() -> 128 bytes; every rank returns the same bytes (torch.distributed broadcast)."""
    import torch
    import torch.distributed as dist

    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = make_id()
        assert len(uid) == 128
        buf = torch.tensor(list(uid), dtype=torch.uint8)
    backend = dist.get_backend(group)
    if backend == "nccl":
        buf = buf.cuda()
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().tolist())


def init_engine_comm(engine, rank, world, group=None):
    """Create the engine's NCCL communicator across the ranks of an initialised torch.distributed job."""
    if world <= 1:
        return
    uid = exchange_unique_id(engine.nccl_unique_id, rank, group)
    engine.comm_init(rank, world, uid)
