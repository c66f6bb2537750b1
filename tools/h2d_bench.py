"""Per-rank pinned host -> device copy bandwidth at N ranks (one process per GPU), plain `copy_(non_blocking=True)` per chunk.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_bench.py [--bind]

Why: the end-to-end leg of bench.py moves 705 MB of y_pred per GPU and step; at N = 1 it runs at the PCIe ceiling, at N = 8 the
per-GPU rate falls to less than half.  This separates the host side (memory placement, shared PCIe switches / root ports)
from our pipeline: every rank copies its own pinned buffer at the same time, nothing else runs.
--bind: pin the process (and, by first touch, its pinned buffer) to the CPUs `nvidia-smi topo` lists for its GPU.
"""
import argparse
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist


def gpu_cpu_affinity(idx):
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        for line in out.splitlines():
            if line.startswith(f"GPU{idx}\t") or line.startswith(f"GPU{idx} "):
                cols = line.split("\t")
                for c in cols:
                    c = c.strip()
                    if c and all(ch.isdigit() or ch in "-," for ch in c) and ("-" in c or "," in c):
                        cpus = set()
                        for part in c.split(","):
                            a, _, b = part.partition("-")
                            cpus.update(range(int(a), int(b or a) + 1))
                        return cpus
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true")
    ap.add_argument("--mb", type=int, default=705)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bound = None
    if args.bind:
        bound = gpu_cpu_affinity(local)
        if bound:
            os.sched_setaffinity(0, bound)
    n = args.mb * (1 << 20) // 4
    host = torch.empty(n, dtype=torch.float32)
    host.fill_(1.0)                        # first touch on the bound CPUs
    host = host.pin_memory()
    devbuf = torch.empty(n, dtype=torch.float32, device="cuda")
    res = {}
    for chunks in (1, 4, 16):
        step = n // chunks
        for two_streams in (False, True):
            streams = [torch.cuda.Stream(), torch.cuda.Stream()] if two_streams else [torch.cuda.current_stream()]
            for _ in range(2):
                for c in range(chunks):
                    with torch.cuda.stream(streams[c % len(streams)]):
                        devbuf[c * step:(c + 1) * step].copy_(host[c * step:(c + 1) * step], non_blocking=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                for c in range(chunks):
                    with torch.cuda.stream(streams[c % len(streams)]):
                        devbuf[c * step:(c + 1) * step].copy_(host[c * step:(c + 1) * step], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            res[(chunks, two_streams)] = reps * step * chunks * 4 / dt / 1e9
    t = torch.tensor([res[k] for k in sorted(res)], device="cuda")
    if world > 1:
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
    else:
        allr = [t]
    if rank == 0:
        print(f"N={world} bind={args.bind} cpus(rank0)={sorted(bound)[:4] if bound else None}... ({len(bound) if bound else os.cpu_count()} cpus)")
        for i, k in enumerate(sorted(res)):
            vals = [float(a[i]) for a in allr]
            print(f"  chunks={k[0]:2d} two_streams={int(k[1])}: per-rank GB/s min {min(vals):6.1f} mean {sum(vals)/len(vals):6.1f} max {max(vals):6.1f}  aggregate {sum(vals):7.1f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
