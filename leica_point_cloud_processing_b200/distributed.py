"""Multi-GPU plumbing: one process per GPU (torchrun), source points sharded by rank, target replicated.

torch.distributed is used only to agree on the NCCL unique id (works over any backend, e.g. gloo on CPU); the
per-evaluation sum of the 14 partial sums over the ranks happens inside the engine (csrc/engine.cu run_cost): fused
into the cost kernel over NVLink peer memory when the ranks could map each other's slot blocks, else one
ncclAllReduce on the engine's own communicator.  No Python sits on the evaluation path.
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


def init_engine_comm(engine, rank, world, group=None, peer_memory=True):
    """Create the engine's NCCL communicator across the ranks of an initialised torch.distributed job and, when the
    GPUs can map each other's memory, switch the per-evaluation sum to the fused peer-memory path.  Returns True
    when the fused path is on."""
    if world <= 1:
        return False
    uid = exchange_unique_id(engine.nccl_unique_id, rank, group)
    engine.comm_init(rank, world, uid)
    if peer_memory and os.environ.get("GICPB_NO_PEER", "0") != "1":
        return enable_peer_reduction(engine, rank, world, group)
    return False


def enable_peer_reduction(engine, rank, world, group=None):
    """Fuse the per-evaluation cross-GPU sum into the cost kernel over NVLink peer memory (include/gicp_b200.h
    gicpb_peer_export / gicpb_peer_import).  Every rank either enables it or none does (an all-gathered vote), so a
    node without peer access simply stays on ncclAllReduce.  Returns True when enabled."""
    import torch.distributed as dist

    try:
        mine = engine.peer_export()
    except Exception:
        mine = None
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    ok = all(h is not None for h in handles)
    if ok:
        try:
            engine.peer_import(handles)
        except Exception:
            ok = False
    votes = [None] * world
    dist.all_gather_object(votes, ok, group=group)
    if not all(votes):
        engine.peer_disable()
        return False
    return True
