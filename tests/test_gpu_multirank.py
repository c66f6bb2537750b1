"""On-hardware multi-rank parity: with the batch sharded over 2 GPUs, the loss partials after the ONE
exchange of the path equal a single GPU working through the same shards - bit for bit in the fp64 vector - and the decode /
render shards need no exchange at all.  Runs `bench.py --check` under torchrun; skipped on boxes with fewer than 2 GPUs
(the driver's 1-GPU test box); `profiles/r2_bench_n2_check.json` / `r2_bench_n8.json` are this check at 2 and 8 GPUs."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_sharded_loss_equals_single_gpu_bitwise():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "3",
                          "--warmup", "3", "--batch", "64", "--check", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["check"] == {"sharded_partials_equal_single_gpu_bitwise": True, "ranks": 2, "images": 128}
    assert line["decode_slow_path_images"] == 0
