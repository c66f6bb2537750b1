"""N > 1 host logic on CPU: two gloo ranks shard a batch contiguously (cvmhot.dist.shard_range), each computes the loss
partials of its shard (the oracle stands in for the kernel here: this test is about the sharding and the one collective,
not about the device math), the fp64 partials vector is all-reduced (cvmhot.dist.allreduce_partials) and finalised
(cvmhot.dist.finalize_partials_host).  The result must equal the single-process batch-global loss — and must NOT equal
the average of the per-shard losses (loss.py:50,59,113,130 normalise by batch-global counts)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from oracle import loss_np, render_np
from oracle.layout import make_layout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _batch(B=6):
    Lo = make_layout(32, 48, 4, "N")
    data = synth.make_batch(Lo, 7, B, ties=False)
    data["boxes"][0] = np.zeros((0, 4))          # an image without objects: object counts differ across shards
    data["cls"][0] = np.zeros((0,), np.int32)
    yt = np.stack([render_np.render_image(Lo, data["boxes"][b], data["cls"][b], data["ignore"][b]) for b in range(B)])
    return Lo, yt, data["y_pred"]


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "computer-vision-models_b200"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from cvmhot import dist as cdist
    r, w, _ = cdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    Lo, yt, yp = _batch()
    lo, hi = cdist.shard_range(yt.shape[0], rank, world)
    part = torch.tensor(loss_np.partials(Lo, yt[lo:hi], yp[lo:hi], True), dtype=torch.float64)
    local_total = loss_np.finalize(Lo, part.tolist())[0]
    cdist.allreduce_partials(part)
    fields = [(f.weight, f.post) for f in Lo.fields]
    q.put((rank, lo, hi, cdist.finalize_partials_host(part.tolist(), fields), local_total))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_covers_everything():
    from cvmhot.dist import shard_range
    for n in (0, 1, 5, 8, 2048):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_loss_equals_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    Lo, yt, yp = _batch()
    ref = loss_np.total_loss(Lo, yt, yp)[0]
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 6)]
    for r in res:
        assert r[3] == pytest.approx(ref, rel=1e-12)          # every rank holds the batch-global loss
    naive = float(np.mean([r[4] for r in res]))
    assert abs(naive - ref) > 1e-6 * abs(ref)                  # averaging per-shard losses would be wrong
