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
            kw = {}
            try:   # the exchange is 128 bytes: its kernel should get the first SM that frees up, not queue behind full-GPU kernels
                kw["pg_options"] = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            except (AttributeError, TypeError):
                pass
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local), **kw)
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_partials(partials, group=None):
    """Sum the loss partials over all ranks (in place) - the one collective of the path.  The ranks' vectors are gathered and
    added in RANK ORDER, so every rank gets the same bits whatever algorithm the backend uses (an all-reduce may add in a
    different order on different ranks or runs)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        gathered = torch.empty((world,) + tuple(partials.shape), dtype=partials.dtype, device=partials.device)
        dist.all_gather_into_tensor(gathered.view(-1), partials.contiguous().view(-1), group=group)
        ordered_sum(gathered, out=partials)
    return partials


def ordered_sum(gathered, out=None):
    """Sum of gathered[r] over r in rank order (sequential adds: bit-reproducible).  On CUDA fp64 [n,16] inputs this is one
    launch of cvm_loss_finalize_gathered; elsewhere a host-ordered loop."""
    if out is None:
        out = torch.empty_like(gathered[0])
    if gathered.is_cuda and gathered.dtype == torch.float64 and gathered.dim() == 2 and gathered.shape[1] == 16:
        from . import ops
        from .layout import Layout
        ops.loss_finalize_gathered(Layout(H=1, W=1, hm=1, nb_classes=1, Cp=1, Ct=2), gathered.contiguous(), partials=out, finalize=False)
        return out
    acc = gathered[0].clone()
    for r in range(1, gathered.shape[0]):
        acc = acc + gathered[r]
    out.copy_(acc)
    return out


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
