"""Per-warp cycle breakdown of decode_scan_kernel (needs a build with EXTRA=-DCVM_DECODE_STATS)."""
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
out = (C.c_ulonglong * 16)()
ops.decode_topk(L, yp, K=100)
lib.cvm_decode_stats(out, 1)
ops.decode_topk(L, yp, K=100)
lib.cvm_decode_stats(out, 1)
v = list(out)
ws = v[6]
names = ["wait cycles", "scan cycles", "test_hits cycles (threshold set)", "tail (release/threshold/gather) cycles", "pixel hits", "test rounds", "warp-steps", "test_hits cycles (no threshold yet)", "append calls", "-", "gather cycles (segment end)", "gather calls (segment end, per warp)", "append: count atomic + shuffles", "append: key / histogram stores", "append: syncwarp + flag vote", "append: rescan"]
for n, x in zip(names, v):
    print(f"{n:45s} {x:14d}  per warp-step {x / max(ws, 1):10.1f}")
print("pixel hits per image", v[4] / B)
