"""Three decode calls at the BASELINE shape (for `ncu --set full -k regex:decode_scan -s 2 -c 1` captures of the third)."""
import sys
sys.path.insert(0, "/root/repo/computer-vision-models_b200"); sys.path.insert(0, "/root/repo")
import torch
from cvmhot import ops
from cvmhot.layout import layout_from_params
from cvmhot.models.centernet import CenternetParams
H, W, C, B = 128, 384, 10, 256
p = CenternetParams(C, True); p.INPUT_HEIGHT, p.INPUT_WIDTH = H * 2, W * 2
L = layout_from_params(p)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
yp = torch.empty((B, H, W, L.Cp), device=dev)
yp[..., :C] = torch.sigmoid(torch.randn((B, H, W, C), device=dev, generator=g) * 1.5 - 4.0)
yp[..., C:] = torch.rand((B, H, W, L.Cp - C), device=dev, generator=g) * 40
for _ in range(3):      # the third call is the steady state (the prediction of the call before is in the workspace)
    out = ops.decode_topk(L, yp, K=100)
torch.cuda.synchronize()
print(float(out["scores"].sum()))
