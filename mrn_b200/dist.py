"""Data-parallel plumbing: one process per GPU, replicated weights, ONE exchange step per iteration -- an NCCL
all-reduce (average) of the trainable-gradient arena over NVLink 5 / NVSwitch.  Replaces the reference's
single-process nn.DataParallel (il_modules/mrn.py:106,133), which re-broadcasts 173 MB of weights and gathers the
full logits to GPU 0 every step.  The loss is computed per shard (samples are independent through experts, router,
gate and CTC), so no activation ever crosses a link.  gloo is used for the CPU tests of the host logic."""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, local_rank, world)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        kw = {"device_id": torch.device("cuda", local_rank)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_bounds(n_items, rank, world):
    """Equal contiguous shards (mean-of-means is exact only for equal shards)."""
    if n_items % world != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (n_items, world))
    per = n_items // world
    return rank * per, (rank + 1) * per


def allreduce_mean_(arena: torch.Tensor):
    """In-place average of a flat gradient arena across ranks (no-op for a single rank)."""
    w = world_size()
    if w > 1:
        if dist.get_backend() == "nccl":
            dist.all_reduce(arena, op=dist.ReduceOp.AVG)      # averaged inside the collective: no separate div_ launch
        else:                                                 # gloo (CPU tests of the host logic) has no AVG
            dist.all_reduce(arena, op=dist.ReduceOp.SUM)
            arena.div_(w)
    return arena


def broadcast_(arena: torch.Tensor, src=0):
    if world_size() > 1:
        dist.broadcast(arena, src=src)
    return arena


def barrier():
    if world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device) -> float:
    if world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
