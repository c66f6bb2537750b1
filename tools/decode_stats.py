"""Per-role cycle breakdown of decode_scan_kernel (needs a build with `make EXTRA=-DCVM_DECODE_STATS`, never the shipped one)."""
import ctypes as C, os, sys
sys.path.insert(0, "/root/repo/computer-vision-models_b200"); sys.path.insert(0, "/root/repo")
import torch
from cvmhot import ops, _lib
from cvmhot.layout import layout_from_params
from cvmhot.models.centernet import CenternetParams
H, W, Cc, B = 128, 384, 10, 256
p = CenternetParams(Cc, True); p.INPUT_HEIGHT, p.INPUT_WIDTH = H * 2, W * 2
L = layout_from_params(p)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
yp = torch.empty((B, H, W, L.Cp), device=dev)
yp[..., :Cc] = torch.sigmoid(torch.randn((B, H, W, Cc), device=dev, generator=g) * 1.5 - 4.0)
yp[..., Cc:] = torch.rand((B, H, W, L.Cp - Cc), device=dev, generator=g) * 40
lib = _lib.lib()
out = (C.c_ulonglong * 32)()
ops.decode_topk(L, yp, K=100)
lib.cvm_decode_stats(out, 1)
ops.decode_topk(L, yp, K=100)
lib.cvm_decode_stats(out, 1)
v = list(out)
n_cta = 148
names = ["scanner: wait full barrier", "scanner: wait first threshold (hint owner)", "scanner: wait segment threshold", "scanner: wait flush",
         "scanner: queue back-pressure (lane cycles)", "scanner: total", "tester: idle", "tester: test_records", "tester: total", "records",
         "batches", "test rounds", "peaks appended", "loader: wait free slot", "loader: total", "tester: gather at segment end", "scanner: wait lookahead granule", "tester: test preamble", "tester: window loads + vote", "tester: append", "scanner: scan (steady)", "scanner: push (steady)"]
for n, x in zip(names, v):
    print(f"{n:48s} {x:16d}  per CTA {x / n_cta:14.1f}")

import numpy as np
ct = (C.c_ulonglong * 1024)()
ops.decode_topk(L, yp, K=100)
lib.cvm_decode_cta_times(ct)
a = np.array(list(ct), dtype=np.int64).reshape(4, 256)[:, :n_cta]
t0 = a[0].min()
for i, nm in enumerate(["start", "loader done", "scanners done", "testers done"]):
    r = (a[i] - t0) / 1000.0
    print(f"{nm:14s} us: min {r.min():8.2f} median {np.median(r):8.2f} max {r.max():8.2f}")
