"""One process per GPU.  Images are independent in render and decode, so the batch is sharded contiguously with no
data-path collective; the loss needs exactly one exchange: the fp64 partials vector (<=16 numbers) is all-reduced
(NCCL on GPUs, gloo in the CPU tests) before the finalise step."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_partials(partials, group=None):
    """Sum the loss partials over all ranks (in place) — the one collective of the path."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def finalize_partials_host(part, fields):
    """Host-side twin of cvm_loss_finalize for the CPU (gloo) tests of the sharding logic: part is a sequence of floats
    [P, N, n_pos, n_obj, field sums...]; fields = [(weight, post), ...]."""
    import math
    P, N, n, nobj = part[0], part[1], part[2], part[3]
    total = (P + N) / n if n > 0 else N
    for i, (weight, post) in enumerate(fields):
        v = part[4 + i] / nobj if nobj > 0 else part[4 + i]
        if post == 1:
            v = math.sqrt(1.0 - 0.99 * math.cos(2.0 * v)) + abs(v * v * 0.05) - 0.0999
        total += v * weight
    return total
