#!/usr/bin/env python
"""Benchmark of the CenterNet / CenterTracker heatmap hot path — BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|4|5] [--batch B] [--check]

One "step" = one pass of the hot path over one batch of synthetic input per GPU:

  --config 2 (default; BASELINE configs[1], and configs[2] at N = 8): render + loss + decode, 256 images per GPU, 10 classes,
             128x384 heatmaps, 32 objects per image, top-K = 100 (11 403 264 algorithmic bytes per image)
  --config 4 (BASELINE configs[3]): CenterTracker, 512 images: previous-frame heatmap render (1 channel, per-object peak)
             + decode with the tracking-offset gather, 16-channel pixels (3 342 336 B per image)
  --config 5 (BASELINE configs[4]): multitask head, 512x1536 maps, 128 images per GPU: CenterNet decode on channels [0:14]
             + semseg argmax on channels [14:19] of the SAME 20-channel tensor in one read (63 700 992 B per image)

Under torchrun (N > 1) every rank owns its own images (weak scaling), the only collective is the exchange of the 16 loss
partials, timing is barrier + device events, max over ranks.  Rank 0 prints ONE JSON line.  --check (N > 1): the
sharded loss partials are compared bit for bit with the same images processed shard by shard on ONE GPU.

`--impl reference` times the reference's own CPU implementation of the same path on the host cores (oracle/ref_harness.py:
the real numba fill_heatmap and to_3channel, the unmodified loss.py over the TF-on-torch shim, and - the reference has no
top-K decode - the NumPy restatement of the canonical decode), one full batch per step where that is affordable.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "computer-vision-models_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

NB_CLASSES, N_OBJ, TOPK = 10, 32, 100
CONFIGS = {
    2: dict(H=128, W=384, batch=256, track=False, wide=0,
            workload="CenterNet 2D-OD full heatmap path (render+loss+decode), 10-class 128x384, 32 obj/img, top-K=100 (BASELINE configs[1])"),
    4: dict(H=128, W=384, batch=512, track=True, wide=0,
            workload="CenterTracker: prev-frame heatmap render + decode with tracking-offset gather, 10-class 128x384, 16-channel pixels, "
                     "32 obj/img, top-K=100 (BASELINE configs[3])"),
    5: dict(H=512, W=1536, batch=128, track=False, wide=20,
            workload="multitask head: CenterNet decode [0:14] + semseg argmax [14:19] of one 20-channel 512x1536 tensor, one read, "
                     "top-K=100 (BASELINE configs[4]: 1024 images over 8 GPUs = 128 per GPU)"),
}
H, W = 128, 384          # (config 2 shape; kept as module attributes for the tests that import gen_objects)
CONFIG_ID = 2


def gen_objects(first_image, B, Hm=H, Wm=W, config_id=CONFIG_ID):
    """per-image rng = default_rng(1234 + 1000*config + image_index); boxes in input px, inside the image."""
    R = 2
    in_w, in_h = Wm * R, Hm * R
    boxes = np.zeros((B, N_OBJ, 4), np.float64)
    cls = np.zeros((B, N_OBJ), np.int32)
    ign = np.zeros((B, 2, 4), np.float64)
    for i in range(B):
        rng = np.random.default_rng(1234 + 1000 * config_id + first_image + i)
        w = np.exp(rng.uniform(np.log(4), np.log(160), N_OBJ))
        h = np.exp(rng.uniform(np.log(4), np.log(96), N_OBJ))
        cx, cy = rng.uniform(0, in_w, N_OBJ), rng.uniform(0, in_h, N_OBJ)
        x0, y0 = np.maximum(0, cx - w / 2), np.maximum(0, cy - h / 2)
        x1, y1 = np.minimum(in_w, cx + w / 2), np.minimum(in_h, cy + h / 2)
        boxes[i] = np.stack([x0, y0, x1 - x0, y1 - y0], axis=1)
        cls[i] = rng.integers(0, NB_CLASSES, N_OBJ)
        ign[i] = np.stack([rng.uniform(0, Wm - 4, 2), rng.uniform(0, Hm - 4, 2), rng.uniform(1, 12, 2), rng.uniform(1, 8, 2)], axis=1)
    return boxes, cls, ign


def centres(boxes, Hm=H, Wm=W):
    cx = np.clip(((boxes[..., 0] + boxes[..., 2] / 2) / 2).astype(np.int64), 0, Wm - 1)
    cy = np.clip(((boxes[..., 1] + boxes[..., 3] / 2) / 2).astype(np.int64), 0, Hm - 1)
    return cx, cy


def prev_records(boxes, first_image):
    """Previous-frame blobs of config 4 (intended behaviour of centertracker/processor.py:22-41 with FN/FP off): one blob per
    object at its (jittered) centre in mask px, size of the box, peak U(0.2, 0.7)."""
    B = boxes.shape[0]
    cx, cy = centres(boxes)
    out = []
    for i in range(B):
        rng = np.random.default_rng(777 + first_image + i)
        jx, jy = rng.integers(-2, 3, N_OBJ), rng.integers(-2, 3, N_OBJ)
        peak = rng.uniform(0.2, 0.7, N_OBJ)
        out.append([(int(cx[i, k] + jx[k]), int(cy[i, k] + jy[k]), float(boxes[i, k, 2]), float(boxes[i, k, 3]), float(peak[k]))
                    for k in range(N_OBJ)])
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def load_traffic(kernel, config_id):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed `ncu --set full` capture
    of this same workload (profiles/traffic.json, written by profiles/ncu_traffic.py); None if not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get(f"config{config_id}", d if config_id == 2 else {}).get(kernel)
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------------
# the reference's CPU implementation of the path (test infrastructure: oracle/), used by --impl reference and cpu_baseline
def _cpu_inputs(config_id, first_image, n_images):
    """Host-side synthetic inputs of `n_images` images of the config (same distributions as the device-side generator)."""
    from oracle.layout import make_layout
    c = CONFIGS[config_id]
    Hm, Wm = c["H"], c["W"]
    Lo = make_layout(Hm, Wm, NB_CLASSES, "N", track=c["track"])
    boxes, cls, ign = gen_objects(first_image, n_images, Hm, Wm, config_id)
    rng = np.random.default_rng(99 + first_image)
    C = c["wide"] or Lo.Cp
    yp = np.empty((n_images, Hm, Wm, C), np.float32)
    yp[..., :NB_CLASSES] = 1.0 / (1.0 + np.exp(-rng.normal(-4.0, 1.5, (n_images, Hm, Wm, NB_CLASSES)).astype(np.float32)))
    yp[..., NB_CLASSES:] = rng.uniform(0, 60, (n_images, Hm, Wm, C - NB_CLASSES)).astype(np.float32)
    return dict(Lo=Lo, boxes=boxes, cls=cls, ign=ign, yp=yp, prev=prev_records(boxes, first_image) if c["track"] else None)


_REF = {}


def _ref():
    if not _REF:
        from oracle import ref_import, ref_harness
        _REF["r"] = ref_import.load()
        _REF["h"] = ref_harness
    return _REF["r"], _REF["h"]


def cpu_render_decode(config_id, inp):
    """Per-image stages of the reference path on one core: returns (seconds dict, y_true or None)."""
    from oracle import decode_np
    r, hz = _ref()
    c = CONFIGS[config_id]
    Lo, n = inp["Lo"], inp["yp"].shape[0]
    t = {}
    yt = None
    t0 = time.perf_counter()
    if config_id == 2:       # REAL fill_heatmap / calc_img_data, loop of processor.py:264-334
        proc, P = hz.make_render_ctx(r, NB_CLASSES, NB_CLASSES, c["H"], c["W"])
        yt = np.stack([hz.render_image_ref(r, proc, P, NB_CLASSES, c["H"], c["W"], inp["boxes"][i], inp["cls"][i], inp["ign"][i])
                       for i in range(n)])
    elif config_id == 4:     # REAL fill_heatmap with explicit centres and peaks (centertracker/processor.py:22-41)
        for i in range(n):
            hm = np.zeros((c["H"], c["W"], 1), np.float32)
            wts = np.ones((c["H"], c["W"]), np.float32)
            for cx, cy, w, h, peak in inp["prev"][i]:
                r["fill_heatmap"](hm, 0.9, 2, wts, cx, cy, w, h, c["W"], c["H"], peak)
    t["render"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    decode_np.decode_topk(Lo, inp["yp"][..., :Lo.Cp], TOPK)      # NumPy restatement (the reference has no top-K decode)
    t["decode"] = time.perf_counter() - t0
    if config_id == 5:       # REAL to_3channel (numba) on the semseg slice
        from numba.typed import List
        items = List([(k, v) for k, v in zip("abcde", [(32, 32, 64), (0, 0, 255), (96, 128, 128), (102, 255, 0), (255, 0, 204)])])
        t0 = time.perf_counter()
        for i in range(n):
            r["to_3channel"](inp["yp"][i, ..., 14:19].copy(), items, None, False, False)
        t["semseg"] = time.perf_counter() - t0
    return t, yt


_WORKER_INPUTS = {}


def _cpu_worker(task):
    config_id, key = task
    return cpu_render_decode(config_id, _WORKER_INPUTS[key])


def cpu_loss(inp, yt):
    """The UNMODIFIED reference loss.py (CenternetLoss.call) executed over the TF-on-torch shim, all host threads."""
    r, hz = _ref()
    loss, _ = hz.ref_loss_object(r, NB_CLASSES, hm=NB_CLASSES)
    t0 = time.perf_counter()
    v = float(loss(yt, inp["yp"]))
    return time.perf_counter() - t0, v


def warm_reference(config_id):
    """numba compilation and imports, outside every timed region."""
    inp = _cpu_inputs(config_id, 0, 1) if config_id != 5 else None
    if inp is not None:
        _, yt = cpu_render_decode(config_id, inp)
        if config_id == 2:
            cpu_loss(inp, yt)
    else:
        r, _ = _ref()
        from numba.typed import List
        items = List([(k, v) for k, v in zip("abcde", [(32, 32, 64), (0, 0, 255), (96, 128, 128), (102, 255, 0), (255, 0, 204)])])
        r["to_3channel"](np.zeros((4, 4, 5), np.float32), items, None, False, False)


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path on ALL host cores.  Render / decode are per image
    (a process pool, one share per core, inputs generated before the fork); the loss needs the whole batch (its counts
    are batch-global) and runs in the parent with torch's intra-op threads.  Under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp
    import torch
    cfg = CONFIGS[args.config]
    cores = max(1, os.cpu_count() or 1)
    torch.set_num_threads(cores)
    # images per step: the full per-GPU batch where a step stays within seconds, else a bounded sample
    n = {2: args.batch or cfg["batch"], 4: min(args.batch or cfg["batch"], 8 * cores), 5: min(args.batch or cfg["batch"], max(2, cores // 4))}[args.config]
    shares = [list(range(i, n, cores)) for i in range(min(cores, n))]
    sizes = [len(s) for s in shares]
    firsts = np.concatenate([[0], np.cumsum(sizes)])[:-1]
    for k, (f, sz) in enumerate(zip(firsts, sizes)):
        _WORKER_INPUTS[k] = _cpu_inputs(args.config, int(f), sz)
    warm_reference(args.config)
    whole_yp = np.concatenate([_WORKER_INPUTS[k]["yp"] for k in range(len(shares))]) if args.config == 2 else None
    stage = {}
    loss_v = None
    with mp.get_context("fork").Pool(len(shares)) as pool:
        tasks = [(args.config, k) for k in range(len(shares))]
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_worker, tasks, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = pool.map(_cpu_worker, tasks, chunksize=1)
            for tt, _ in res:
                for k, v in tt.items():
                    stage[k] = stage.get(k, 0.0) + v
            if args.config == 2:
                yt = np.concatenate([r_[1] for r_ in res])
                dt, loss_v = cpu_loss(dict(yp=whole_yp), yt)
                stage["loss_wall"] = stage.get("loss_wall", 0.0) + dt
        tot = time.perf_counter() - t0
    value = n * args.steps / tot
    same = n == (args.batch or cfg["batch"])
    kinds = {2: "reference (real numba fill_heatmap + unmodified loss.py over the TF-on-torch shim) + port (NumPy top-K decode: the reference has none)",
             4: "reference (real numba fill_heatmap) + port (NumPy top-K decode with track gather)",
             5: "reference (real numba to_3channel) + port (NumPy top-K decode)"}[args.config]
    line = {
        "impl": "reference", "metric": "images/sec (heatmap render+loss+decode)", "value": value, "unit": "images/sec",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "batch_per_step": n,
                   "note": "the full per-GPU batch per step" if same else "bounded sample of the per-GPU batch (same per-image work)"},
        "cpu_baseline": {"value": value, "unit": "images/sec", "cores": cores, "kind": "reference",
                         "sample": f"{n} images/step x {args.steps} steps on {cores} host cores; {kinds}"},
        "e2e": {"value": value, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stages_core_s_per_image": {k: float(v / (n * args.steps)) for k, v in stage.items()},
        "loss": loss_v,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def make_pred(torch, dev, g, B, Hm, Wm, L, C_total, boxes, cls):
    """Synthetic y_pred on the device: sigmoid-noise heatmaps with the object centres raised, heads filled."""
    y_pred = torch.empty((B, Hm, Wm, C_total), dtype=torch.float32, device=dev)
    chunk = max(1, min(B, (1 << 28) // (Hm * Wm * C_total)))       # bounded temporaries for the large maps
    for a in range(0, B, chunk):
        b = min(B, a + chunk)
        y_pred[a:b, ..., :L.hm] = torch.sigmoid(torch.randn((b - a, Hm, Wm, L.hm), device=dev, generator=g) * 1.5 - 4.0)
        y_pred[a:b, ..., L.off_roff:L.off_roff + 2] = torch.rand((b - a, Hm, Wm, 2), device=dev, generator=g)
        y_pred[a:b, ..., L.off_box:L.off_box + 2] = torch.rand((b - a, Hm, Wm, 2), device=dev, generator=g) * 120 + 4
        if L.off_track >= 0:
            y_pred[a:b, ..., L.off_track:L.off_track + 2] = torch.randn((b - a, Hm, Wm, 2), device=dev, generator=g) * 8
        if C_total > L.Cp:
            y_pred[a:b, ..., L.Cp:] = torch.randn((b - a, Hm, Wm, C_total - L.Cp), device=dev, generator=g)
    cx, cy = centres(boxes, Hm, Wm)
    bi = np.repeat(np.arange(B), N_OBJ)
    peak = torch.rand(B * N_OBJ, device=dev, generator=g) * 0.69 + 0.3
    y_pred[torch.from_numpy(bi).to(dev), torch.from_numpy(cy.reshape(-1)).to(dev), torch.from_numpy(cx.reshape(-1)).to(dev),
           torch.from_numpy(cls.reshape(-1).astype(np.int64)).to(dev)] = peak
    return y_pred


def run_ours(args):
    """Stream priorities (measured, B200): with the render on the main stream (configs[3], --overlap 0 / 1) the main stream
    is HIGH priority and the decode's stream default - the block scheduler gives render / loss CTAs the SMs first and the
    decode fills in behind them.  With the three-stream schedule of configs[1] (--overlap 2) the RENDER stream is the high-
    priority one (0.491 against 0.496 ms per step with the loss stream high): the next batch's ground truth is the work
    that must never wait, the loss of the current batch takes what it frees."""
    import torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    piped = args.config == 2 and args.overlap == 2
    with torch.cuda.stream(torch.cuda.Stream(device=torch.device("cuda", local), priority=0 if piped else -1)):
        _run_ours(args)


def _run_ours(args):
    import torch
    import torch.distributed as dist
    from cvmhot import dist as cdist
    from cvmhot import ops
    from cvmhot.layout import layout_from_params
    from cvmhot.models.centernet import CenternetParams
    from cvmhot.models.centertracker import CentertrackerParams, CenterTrackerProcess
    from cvmhot.models.centernet.processor import pack_boxes, pack_objects

    rank, world, local = cdist.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = CONFIGS[args.config]
    Hm, Wm = cfg["H"], cfg["W"]
    B = args.batch or cfg["batch"]
    ops.set_decode_spare_sms(args.spare_sms if args.spare_sms >= 0 else (2 if world > 1 and args.config == 2 else 0))

    p = (CentertrackerParams if cfg["track"] else CenternetParams)(NB_CLASSES, per_class_heatmap=True)
    p.INPUT_HEIGHT, p.INPUT_WIDTH = Hm * 2, Wm * 2
    L = layout_from_params(p)
    C_total = cfg["wide"] or L.Cp
    assert L.Cp == (16 if cfg["track"] else 14)

    # ---- synthetic inputs, resident in HBM before the timed region ----
    def device_inputs(first_image, seed):
        boxes, cls, ign = gen_objects(first_image, B, Hm, Wm, args.config)
        g = torch.Generator(device=dev).manual_seed(seed)
        d = dict(boxes=boxes, cls=cls, ign=ign, y_pred=make_pred(torch, dev, g, B, Hm, Wm, L, C_total, boxes, cls))
        if args.config == 2:
            d["rec"], d["offs"] = pack_objects(list(boxes), list(cls))
            d["ign_rec"], d["ign_offs"] = pack_boxes(list(ign))
        if args.config == 4:
            d["rec"], d["offs"] = CenterTrackerProcess.pack_prev_records(prev_records(boxes, first_image))
        if "rec" in d:
            d["objs_d"] = ops.to_device_records(d["rec"], ops.OBJ_DTYPE, dev)
            d["offs_d"] = torch.from_numpy(d["offs"]).to(dev)
        if "ign_rec" in d:
            d["ign_d"] = ops.to_device_records(d["ign_rec"], ops.BOX_DTYPE, dev)
            d["ioffs_d"] = torch.from_numpy(d["ign_offs"]).to(dev)
        return d

    inp = device_inputs(rank * B, 1234 + rank)
    y_pred = inp["y_pred"]
    y_pred_cn = y_pred[..., :L.Cp]                       # (config 5: the CenterNet slice of the wide tensor, read in place)
    y_true = torch.empty((B, Hm, Wm, L.Ct), dtype=torch.float32, device=dev) if args.config == 2 else None
    prev_hm = torch.empty((B, Hm, Wm, 1), dtype=torch.float32, device=dev) if args.config == 4 else None
    partials = torch.empty(16, dtype=torch.float64, device=dev)
    gathered = torch.empty((world, 16), dtype=torch.float64, device=dev)
    partials_sum = torch.zeros(16, dtype=torch.float64, device=dev)     # N > 1: the rank-ordered sum of all ranks' partials
    loss_out = torch.zeros(10, dtype=torch.float32, device=dev)
    px = Hm * Wm
    if args.config == 2:
        names = ["render_kernel", "loss_fwd_fast_kernel", "decode_scan_kernel"]
        sbytes = [4 * px * L.Ct * B, 4 * px * (L.Ct + L.Cp) * B, 4 * px * L.Cp * B]
        launches = 4 if world == 1 else 5      # render, loss (+ reduce + finalize in its last block), scan, merge (+ finalize at N > 1)
    elif args.config == 4:
        names = ["render_kernel(prev_hm)", "decode_scan_kernel"]
        sbytes = [4 * px * B, 4 * px * L.Cp * B]
        launches = 3
    else:
        names = ["decode_scan_kernel(+semseg argmax)"]
        sbytes = [(4 * px * C_total + px) * B]
        launches = 2

    ev = lambda: torch.cuda.Event(enable_timing=True)

    side = torch.cuda.Stream(device=dev)
    comm = torch.cuda.Stream(device=dev)      # N > 1: the exchange of the loss partials and the finalise step

    def step(events=None, ov=False):
        """One pass of the path.  ov: the decode - it only reads y_pred - is issued on a second stream beside render (+ loss).
        All kernels are persistent and fill the GPU, so they still execute one after the other, but the CTAs of the next
        kernel start on the SMs the previous one frees instead of waiting for its last CTA (each kernel's start-up and tail
        cost ~10 us at these batch sizes).  events (sequential schedule only): one mark per stage boundary."""
        k = 0

        def mark():
            nonlocal k
            if events is not None:
                events[k].record()
            k += 1

        def decode():
            if args.config == 5:
                return ops.decode_topk(L, y_pred_cn, K=TOPK, semseg=(14, 5))
            return ops.decode_topk(L, y_pred, K=TOPK)

        main = torch.cuda.current_stream(dev)
        out = None
        if ov:
            side.wait_stream(main)     # (not before the previous step's render + loss are through)

        def decode_beside():
            # issued AFTER render / loss: with the main stream at high priority the decode's CTAs take the SMs those free
            with torch.cuda.stream(side):
                return decode()
        mark()
        if args.config == 2:
            piped = ov and args.overlap == 2
            if piped:
                # the render has its own stream and two y_true buffers: render n + 1 (the next batch's ground truth, as the
                # reference's generator workers prepare it while the model trains) fills the SMs that loss n frees
                i = step_no[0] % 2
                step_no[0] += 1
                yt = y_true_pair[i]
                with torch.cuda.stream(rstream):
                    if loss_done[i] is not None:
                        rstream.wait_event(loss_done[i])      # the loss that read this buffer two steps ago
                    ops.render_gt(L, inp["objs_d"], inp["offs_d"], B, inp["ign_d"], inp["ioffs_d"], out=yt)
                    rendered = torch.cuda.Event()
                    rendered.record(rstream)
                main.wait_event(rendered)
            else:
                yt = y_true
                ops.render_gt(L, inp["objs_d"], inp["offs_d"], B, inp["ign_d"], inp["ioffs_d"], out=yt)
            mark()
            if world == 1:
                ops.loss_total(L, yt, y_pred, True, partials=partials, out=loss_out)     # one launch
                if piped:
                    loss_done[i] = torch.cuda.Event()
                    loss_done[i].record(main)
                mark()
                out = decode_beside() if ov else decode()
            else:
                # the one collective of the path: 16 doubles per rank, gathered while the decode kernel runs and summed in
                # rank order (bit-reproducible whatever the collective's algorithm)
                if ov:
                    # exchange + finalise on their own stream: the main stream (render, loss) never waits for the other
                    # ranks, so rank skew is absorbed with a step of slack instead of being paid every step
                    j = xchg_no[0] % 2                         # two send / receive buffers: the loss never waits for the
                    xchg_no[0] += 1                            # exchange of the step before, only for the one before that
                    if gather_done[j] is not None:
                        main.wait_event(gather_done[j])
                    ops.loss_partials(L, yt, y_pred, True, out=partials_pair[j])
                    ready = torch.cuda.Event()
                    ready.record(main)
                    if piped:
                        loss_done[i] = ready
                    with torch.cuda.stream(comm):
                        comm.wait_event(ready)
                        dist.all_gather_into_tensor(gathered_pair[j].view(-1), partials_pair[j])
                        ops.loss_finalize_gathered(L, gathered_pair[j], partials=partials_sum, out=loss_out)     # rank-ordered sum + finalise: one launch
                        gather_done[j] = torch.cuda.Event()
                        gather_done[j].record(comm)
                    mark()
                    out = decode_beside()
                else:
                    ops.loss_partials(L, yt, y_pred, True, out=partials)
                    work = dist.all_gather_into_tensor(gathered.view(-1), partials, async_op=True)
                    mark()
                    out = decode()
                    work.wait()
                    ops.loss_finalize_gathered(L, gathered, partials=partials_sum, out=loss_out)
        elif args.config == 4:
            ops.render_prev_heatmap(L, inp["objs_d"], inp["offs_d"], B, out=prev_hm)
            mark()
            out = decode_beside() if ov else decode()
        else:
            out = decode()
        mark()
        if ov:
            # the main stream may run ONE step ahead of the decode stream: render n + 1 takes the SMs that decode n frees
            # (no data flows between them; the decode only reads y_pred).  join() closes the pipeline.
            done = torch.cuda.Event()
            done.record(side)
            if decode_done:
                main.wait_event(decode_done.pop())
            decode_done.append(done)
        return out

    decode_done, gather_done, xchg_no = [], [None, None], [0]
    partials_pair = [partials, torch.empty_like(partials)]
    gathered_pair = [gathered, torch.empty_like(gathered)]
    step_no, loss_done = [0], [None, None]
    rstream = torch.cuda.Stream(device=dev, priority=-1)      # --overlap 2: the render stream (see run_ours)
    y_true_pair = [y_true, torch.empty_like(y_true)] if (args.overlap == 2 and args.config == 2) else None

    def join():
        main = torch.cuda.current_stream(dev)
        main.wait_stream(side)
        main.wait_stream(comm)
        main.wait_stream(rstream)
        decode_done.clear()
        gather_done[0] = gather_done[1] = None

    n_marks = len(names) + 1
    overlap = bool(args.overlap) and args.config in (2, 4)
    for _ in range(max(args.warmup, 3)):
        out = step(ov=overlap)
    join()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: device-resident inputs (`value`) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    t_host = time.perf_counter()
    for k in range(args.steps):
        out = step(ov=overlap)
    host_issue_ms = (time.perf_counter() - t_host) * 1e3 / args.steps     # Python + launch cost per step (no sync inside)
    join()        # every decode of the timed steps has finished before the closing event
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)

    # ---- per-stage durations (roofline attribution): the same step in the sequential schedule, one event per stage boundary.
    #      (With the decode on its own stream the stages interleave and cannot be told apart; `value` is the overlapped run.) ----
    n_stage_steps = max(3, min(args.steps, 10))
    stage_events = [[ev() for _ in range(n_marks)] for _ in range(n_stage_steps)]
    for _ in range(2):
        step()
    barrier()
    s0, s1 = ev(), ev()
    s0.record()
    for k in range(n_stage_steps):
        out = step(stage_events[k])
    s1.record()
    barrier()
    sequential_ms = s0.elapsed_time(s1) / n_stage_steps
    # (median over the steps: a host-side hiccup - allocator, garbage collection - between two launches of one step would
    # otherwise be booked on whichever stage was waiting for its launch)
    stage_ms = np.median(np.array([[s[i].elapsed_time(s[i + 1]) for i in range(n_marks - 1)] for s in stage_events]), axis=0)
    sequential_ms = float(np.median([s[0].elapsed_time(s[n_marks - 1]) for s in stage_events]))     # (median, for the same reason)

    # ---- the backward of the loss as its own timed stage (config 2; not part of the metric: reads y_true + y_pred, writes the gradient) ----
    extra = {}
    if args.config == 2:
        grad = torch.empty_like(y_pred)
        for _ in range(3):
            ops.loss_backward(L, y_true, y_pred, partials_sum if world > 1 else partials, out=grad)
        b0, b1 = ev(), ev()
        torch.cuda.synchronize()
        b0.record()
        for _ in range(args.steps):
            ops.loss_backward(L, y_true, y_pred, partials_sum if world > 1 else partials, out=grad)
        b1.record()
        torch.cuda.synchronize()
        bwd_ms = b0.elapsed_time(b1) / args.steps
        bwd_bytes = 4 * px * (L.Ct + 2 * L.Cp) * B
        extra["loss_bwd_fast_kernel"] = {"ms": bwd_ms, "bytes": bwd_bytes}
        del grad

    # ---- --check: the sharded partials equal the same images processed shard by shard on one GPU, bit for bit ----
    check = None
    if args.check and args.config == 2:
        mine = (partials_sum if world > 1 else partials).clone()
        shard_parts = []
        for r_ in range(world):
            o = inp if r_ == rank else device_inputs(r_ * B, 1234 + r_)
            yt_r = torch.empty_like(y_true)
            ops.render_gt(L, o["objs_d"], o["offs_d"], B, o["ign_d"], o["ioffs_d"], out=yt_r)
            shard_parts.append(ops.loss_partials(L, yt_r, o["y_pred"], True).clone())
            del yt_r
            if r_ != rank:
                del o
        single = cdist.ordered_sum(torch.stack(shard_parts))
        ok = torch.tensor([1.0 if torch.equal(single, mine) else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        check = {"sharded_partials_equal_single_gpu_bitwise": bool(ok.item() == 1.0), "ranks": world, "images": world * B}

    # ---- e2e: the same step through the public API with HOST buffers (pinned), copies inside the timed region ----
    e2e_steps = max(2, min(args.steps, 5 if args.config != 5 else 2))
    res_h = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items() if v is not None}
    loss_h = torch.empty(1, dtype=torch.float32).pin_memory()
    y_pred_h = torch.empty(y_pred.shape, dtype=torch.float32).pin_memory()
    y_pred_h.copy_(y_pred)
    small_h = {k: torch.from_numpy(v.view(np.uint8).reshape(-1) if v.dtype.fields else v).pin_memory()
               for k, v in inp.items() if k in ("rec", "offs", "ign_rec", "ign_offs")}
    h2d = y_pred_h.numel() * 4 + sum(v.numel() * v.element_size() for v in small_h.values())
    d2h = sum(v.numel() * v.element_size() for v in res_h.values()) + (loss_h.numel() * 4 if args.config == 2 else 0)
    y_pred_in = torch.empty_like(y_pred)

    # y_pred dominates the step (PCIe), so the batch is cut into chunks whose host->device copies run on a copy stream while
    # the previous chunk is in the kernels: loss partials are additive over chunks, render and decode are per image.
    n_chunks = args.chunks
    bounds = [(k * B // n_chunks, (k + 1) * B // n_chunks) for k in range(n_chunks)]
    copy_stream = torch.cuda.Stream(device=dev)
    # the e2e leg goes through the reference-signature mirror (the calls a user of the reference makes): ProcessImages
    # (render), CenternetLoss.call (loss; with process_group under torchrun), decode_topk(y_pred, params) (decode)
    from cvmhot.models.centernet import CenternetLoss, ProcessImages
    from cvmhot.models.centernet.post_processing import decode_topk as mirror_decode_topk
    proc = ProcessImages(p, device=dev) if args.config == 2 else None
    mirror_loss = CenternetLoss(p, process_group=dist.group.WORLD if world > 1 else None) if args.config == 2 else None

    def e2e_step():
        main = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(main)          # y_pred_in of the previous step has been consumed
        ready = []
        with torch.cuda.stream(copy_stream):
            for a, b in bounds:
                y_pred_in[a:b].copy_(y_pred_h[a:b], non_blocking=True)
                e = torch.cuda.Event()
                e.record(copy_stream)
                ready.append(e)
        if args.config == 2:      # host records -> device -> y_true (ProcessImages.render_packed copies them itself)
            proc.render_packed(L, inp["rec"], inp["offs"], inp["ign_rec"], inp["ign_offs"], out=y_true)
        elif args.config == 4:
            sd = {k: v.to(dev, non_blocking=True) for k, v in small_h.items()}
            ops.render_prev_heatmap(L, sd["rec"], sd["offs"], B, out=prev_hm)
        outs = []
        for k, (a, b) in enumerate(bounds):
            main.wait_event(ready[k])
            if args.config == 5:
                outs.append(ops.decode_topk(L, y_pred_in[a:b, ..., :L.Cp], K=TOPK, semseg=(14, 5)))
            else:
                outs.append(mirror_decode_topk(y_pred_in[a:b], p, K=TOPK))
        if args.config == 2:      # the loss is batch-global: one call once the last chunk has landed (0.26 ms against 13 ms of copies)
            loss_h.copy_(mirror_loss(y_true, y_pred_in).detach().reshape(1), non_blocking=True)
        for k, (a, b) in enumerate(bounds):
            for k_, v in res_h.items():
                v[a:b].copy_(outs[k][k_], non_blocking=True)

    e2e_step()
    barrier()
    f0, f1 = ev(), ev()
    f0.record()
    for _ in range(e2e_steps):
        e2e_step()
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (elapsed_ms * 1e-3)
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    peak_gbs, peak_src = load_peaks()
    dom = int(np.argmax(stage_ms))
    achieved = sbytes[dom] / (stage_ms[dom] * 1e-3) / 1e9
    whole = sum(sbytes) / (elapsed_ms / args.steps * 1e-3) / 1e9
    stages = {n: {"ms": float(m), "gbs": b / (m * 1e-3) / 1e9, "frac": b / (m * 1e-3) / 1e9 / peak_gbs}
              for n, m, b in zip(names, stage_ms, sbytes)}
    for n, v in extra.items():
        stages[n] = {"ms": v["ms"], "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9, "frac": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak_gbs,
                     "note": "timed on its own after the step loop; not part of `value`"}

    # CPU baseline (reported only): the reference path on a bounded sample, ONE core, rank 0, N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = {2: 8, 4: 8, 5: 1}[args.config]
        warm_reference(args.config)
        ci = _cpu_inputs(args.config, 0, n_cpu)
        import torch as _t
        nt = _t.get_num_threads()
        _t.set_num_threads(1)
        ts, yt = cpu_render_decode(args.config, ci)
        if args.config == 2:
            ts["loss"] = cpu_loss(ci, yt)[0]
        _t.set_num_threads(nt)
        cpu = {"value": n_cpu / sum(ts.values()), "unit": "images/sec", "cores": 1, "kind": "reference",
               "sample": f"{n_cpu} images of the same workload; real numba fill_heatmap / to_3channel, the unmodified loss.py over the "
                         f"TF-on-torch shim, NumPy top-K decode (the reference has none); " + " ".join(f"{k} {v:.2f}s" for k, v in ts.items())}

    line = {
        "metric": "images/sec (heatmap render+loss+decode)", "value": value, "unit": "images/sec", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world} (batch shards" + (", one 128-byte exchange overlapped with decode)" if args.config == 2 else ", no collective)"),
                   "l2": f"inputs larger than L2 ({y_pred.numel() * 4 / 1e6:.0f} MB y_pred per GPU per step)",
                   "schedule": ((("render of batch n+1 on its own stream (two y_true buffers) beside loss n; " if (args.overlap == 2 and args.config == 2) else "")
                                 + "decode on a second stream (it only reads y_pred); every step's render, loss and decode complete "
                                 "inside the timed region; per-stage times from a "
                                 f"separate sequential pass of the same step ({sequential_ms:.4f} ms/step)")) if overlap else "one stream, stages back to back"},
        "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": load_traffic(names[dom].split("(")[0], args.config), "peak_source": peak_src,
                     "whole_step_gbs": whole, "whole_step_frac": whole / peak_gbs, "stages": stages},
        "host_issue_ms_per_step": host_issue_ms,
        "sequential": {"ms_per_step": sequential_ms, "value": B * world / (sequential_ms * 1e-3),
                       "note": "the same step with all stages back to back on one stream (the pass the stage times come from)"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "note": f"host pinned buffers -> the reference-signature mirror (ProcessImages / CenternetLoss / decode_topk) -> host results, copies inside the timed region "
                                             f"({n_chunks} chunks: H2D of chunk k+1 overlaps the kernels of chunk k)"},
        "gpu_launches": launches * args.steps,
        "clocks": clocks,
        "decode_slow_path_images": int(ops.decode_fallback_count()),
    }
    if args.config == 2:
        line["loss"] = float(loss_out[0])
    if check is not None:
        line["check"] = check
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5], help="BASELINE.json configs[config-1]")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the config's)")
    ap.add_argument("--chunks", type=int, default=4, help="e2e leg: H2D chunks per step")
    ap.add_argument("--spare-sms", type=int, default=-1, help="SMs the decode leaves to the collective (default: 2 when N > 1)")
    ap.add_argument("--overlap", type=int, default=2, help="configs 2 / 4: 0 = one stream; 1 = decode on a second stream beside render (+ loss); "
                    "2 (config 2) = also the render on its own stream with two y_true buffers, one batch ahead of the loss")
    ap.add_argument("--check", action="store_true", help="N > 1: compare the sharded loss partials with one GPU, bit for bit")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
