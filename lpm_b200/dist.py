"""One-process-per-GPU plumbing: torch.distributed carries the NCCL unique id to all ranks, the engine's
own NCCL communicator (bound inside liblpmx.so) then moves the per-stage source records over NVLink."""
import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_unique_id(make_id, rank):
    """Rank 0 creates the id (make_id()), everyone receives the same 128 bytes via torch.distributed
    (works on gloo and nccl process groups)."""
    import torch.distributed as dist
    obj = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]


def init_engine_comm(engine, rank, world):
    """Give `engine` a NCCL communicator spanning the torch.distributed world and set its partition."""
    from .api import Engine
    if world == 1:
        engine.set_partition(0, 1)
        return
    uid = broadcast_unique_id(Engine.comm_unique_id, rank)
    engine.comm_init(uid, rank, world)
