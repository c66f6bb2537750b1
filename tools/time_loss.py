import os, sys
sys.path.insert(0, "/root/repo/computer-vision-models_b200"); sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from cvmhot import ops
from cvmhot.layout import layout_from_params
from cvmhot.models.centernet import CenternetParams
from cvmhot.models.centernet.processor import pack_boxes, pack_objects
B = 256
p = CenternetParams(10, per_class_heatmap=True); p.INPUT_HEIGHT, p.INPUT_WIDTH = 256, 768
L = layout_from_params(p)
dev = torch.device("cuda", 0)
boxes, cls, ign = bench.gen_objects(0, B)
rec, offs = pack_objects(list(boxes), list(cls)); ign_rec, ign_offs = pack_boxes(list(ign))
y_true = ops.render_gt(L, ops.to_device_records(rec, ops.OBJ_DTYPE, dev), torch.from_numpy(offs).to(dev), B,
                       ops.to_device_records(ign_rec, ops.BOX_DTYPE, dev), torch.from_numpy(ign_offs).to(dev))
g = torch.Generator(device=dev).manual_seed(1)
yp = torch.empty((B, 128, 384, L.Cp), device=dev)
yp[..., :10] = torch.sigmoid(torch.randn((B, 128, 384, 10), device=dev, generator=g) * 1.5 - 4.0)
yp[..., 10:] = torch.rand((B, 128, 384, L.Cp - 10), device=dev, generator=g) * 40
yt2, yp2 = y_true.clone(), yp.clone()
part = torch.empty(16, dtype=torch.float64, device=dev)
_queue = torch.empty(2**27, device=dev)
for _ in range(3): ops.loss_partials(L, y_true, yp, True, out=part)
torch.cuda.synchronize()
for _ in range(30): _queue.fill_(1.0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for i in range(n): ops.loss_partials(L, y_true if i % 2 else yt2, yp if i % 2 else yp2, True, out=part)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{os.environ.get('CVM_TAG', ''):12s} {ms:.4f} ms  {(y_true.numel() + yp.numel()) * 4 / ms / 1e6:.0f} GB/s  loss partial[1]={float(part[1]):.6f}", flush=True)

# backward (d total / d y_pred): reads y_true + y_pred, writes grad
grad = torch.empty_like(yp)
for _ in range(3): ops.loss_backward(L, y_true, yp, part, out=grad)
torch.cuda.synchronize()
for _ in range(30): _queue.fill_(1.0)
e0.record()
for i in range(n): ops.loss_backward(L, y_true if i % 2 else yt2, yp if i % 2 else yp2, part, out=grad)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"backward     {ms:.4f} ms  {(y_true.numel() + 2 * yp.numel()) * 4 / ms / 1e6:.0f} GB/s", flush=True)
