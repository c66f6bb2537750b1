import os, sys
sys.path.insert(0, "/root/repo/computer-vision-models_b200"); sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from cvmhot import ops
from cvmhot.layout import layout_from_params
from cvmhot.models.centernet import CenternetParams
from cvmhot.models.centernet.processor import pack_boxes, pack_objects
B = int(os.environ.get("CVM_B", "256"))
p = CenternetParams(10, per_class_heatmap=True); p.INPUT_HEIGHT, p.INPUT_WIDTH = 256, 768
L = layout_from_params(p)
dev = torch.device("cuda", 0)
boxes, cls, ign = bench.gen_objects(0, B)
rec, offs = pack_objects(list(boxes), list(cls)); ign_rec, ign_offs = pack_boxes(list(ign))
objs_d = ops.to_device_records(rec, ops.OBJ_DTYPE, dev); offs_d = torch.from_numpy(offs).to(dev)
ign_d = ops.to_device_records(ign_rec, ops.BOX_DTYPE, dev); ioffs_d = torch.from_numpy(ign_offs).to(dev)
y1 = torch.empty((B, 128, 384, L.Ct), device=dev); y2 = torch.empty_like(y1)
_queue = torch.empty(2**27, device=dev)


def run(tag):
    for _ in range(3): ops.render_gt(L, objs_d, offs_d, B, ign_d, ioffs_d, out=y1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    for _ in range(30): _queue.fill_(1.0)      # ~5 ms of queued GPU work: the timed launches below are enqueued while the GPU
    e0.record()                                # is still busy, so the events see kernel time, not Python launch overhead
    for i in range(n): ops.render_gt(L, objs_d, offs_d, B, ign_d, ioffs_d, out=y1 if i % 2 else y2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{tag:30s} {ms:.4f} ms  {y1.numel()*4/ms/1e6:.0f} GB/s", flush=True)
for env in sys.argv[1:]:
    for kv in env.split(","):
        if kv != "-":
            k, v = kv.split("="); os.environ[k] = v
    run(env)
    for kv in env.split(","):
        if kv != "-": os.environ.pop(kv.split("=")[0])
if os.environ.get("CVM_RENDER_DBG"):
    import ctypes
    from cvmhot import _lib
    lib = _lib.lib()
    out = (ctypes.c_ulonglong * 12)()
    lib.cvm_render_debug(out)
    ops.render_gt(L, objs_d, offs_d, B, ign_d, ioffs_d, out=y1); torch.cuda.synchronize()
    lib.cvm_render_debug(out)
    names = ["loop+wait_read", "barC", "fill", "poll", "barA", "scatter", "ignore+fence", "barB", "ballot", "n_objects", "splat_objects", "-"]
    tot = sum(out)
    print({n: int(v / (148 * 4)) for n, v in zip(names, out)}, "cycles per group leader; total", int(tot / 592))
