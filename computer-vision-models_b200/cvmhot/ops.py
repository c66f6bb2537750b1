"""Tensor-level entry points: torch CUDA tensors in/out, all math inside libcvmhot.so (C ABI, include/cvmhot.h).

torch is only used for device memory and streams.  Every function enqueues on torch's current CUDA stream and never
synchronises.  There is no CPU path: passing a CPU tensor raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .layout import Layout

OBJ_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("w", "<f8"), ("h", "<f8"), ("cx", "<i4"), ("cy", "<i4"),
                      ("cls", "<i4"), ("flags", "<i4"), ("peak", "<f4"), ("track", "<f4", (2,)), ("_pad", "<f4")])
BOX_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("w", "<f8"), ("h", "<f8")])
ROI_DTYPE = np.dtype([("inv_scale", "<f4"), ("off_left", "<f4"), ("off_top", "<f4"), ("_pad", "<f4")])
assert OBJ_DTYPE.itemsize == 64 and BOX_DTYPE.itemsize == 32 and ROI_DTYPE.itemsize == 16


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.CvmError(f"{name} must be a CUDA tensor (there is no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise _lib.CvmError(f"{name} must be {dtype}, got {t.dtype}")


def _pixel_strided(t, name):
    """Accept [..., C] tensors that are contiguous up to a channel slice; return the per-pixel stride in floats."""
    _need_cuda(t, name, torch.float32)
    if t.dim() < 2 or t.stride(-1) != 1:
        raise _lib.CvmError(f"{name}: channel dimension must be innermost and dense")
    st = t.stride(-2)
    expect = st
    for d in range(t.dim() - 2, -1, -1):
        if t.size(d) != 1 and t.stride(d) != expect:
            raise _lib.CvmError(f"{name}: only NHWC-contiguous tensors (or channel slices of them) are supported")
        expect *= t.size(d)
    return int(st)


def to_device_records(arr, dtype, device):
    """numpy structured array -> uint8 CUDA tensor holding the same bytes."""
    arr = np.ascontiguousarray(arr, dtype=dtype)
    if arr.size == 0:
        return torch.zeros(dtype.itemsize, dtype=torch.uint8, device=device)   # never dereferenced
    return torch.from_numpy(arr.view(np.uint8).reshape(-1)).to(device, non_blocking=True)


def make_rois(rois, device):
    """rois: iterable of (scale, offset_left, offset_top) per image -> device records (inv_scale rounded like NumPy does)."""
    a = np.zeros(len(rois), dtype=ROI_DTYPE)
    for i, (scale, ol, ot) in enumerate(rois):
        a[i]["inv_scale"] = np.float32(1.0 / scale)
        a[i]["off_left"] = np.float32(ol)
        a[i]["off_top"] = np.float32(ot)
    return to_device_records(a, ROI_DTYPE, device)


_ws_cache = {}


def _workspace(device, nbytes, tag):
    # one scratch buffer per (device, stream, purpose): concurrent callers on different streams never share it
    key = (str(device), torch.cuda.current_stream(device).cuda_stream, tag)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


# ---- pre-processing front-end ----------------------------------------------------------------------------------------
def prepare_objects(raw_boxes, raw_cls, raw_offsets, img_w, img_h, min_box_area, raw_track=None):
    """Device-side clip_to_img + MIN_BOX_AREA filter (reference processor.py:46-56,241-253) for a whole batch.

    raw_boxes [n,4] float64, raw_cls [n] int32, raw_offsets [B+1] int32 (all CUDA), raw_track [n,2] float32 or None.
    Returns (objs, obj_offsets, ignore, ign_offsets): uint8 record tensors + int32 offsets, ready for render_gt."""
    _need_cuda(raw_boxes, "raw_boxes", torch.float64)
    _need_cuda(raw_cls, "raw_cls", torch.int32)
    _need_cuda(raw_offsets, "raw_offsets", torch.int32)
    if raw_track is not None:
        _need_cuda(raw_track, "raw_track", torch.float32)
    dev = raw_boxes.device
    n, B = int(raw_cls.numel()), int(raw_offsets.numel()) - 1
    if raw_boxes.numel() != 4 * n or not raw_boxes.is_contiguous():
        raise _lib.CvmError("raw_boxes must be a contiguous [n,4] float64 tensor")
    objs = torch.empty(max(n, 1) * OBJ_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    ignore = torch.empty(max(n, 1) * BOX_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    obj_offsets = torch.empty(B + 1, dtype=torch.int32, device=dev)
    ign_offsets = torch.empty(B + 1, dtype=torch.int32, device=dev)
    rc = _lib.lib().cvm_prepare_objects(_ptr(raw_boxes), _ptr(raw_cls), _ptr(raw_track), _ptr(raw_offsets), B, float(img_w),
                                        float(img_h), float(min_box_area), _ptr(objs), _ptr(obj_offsets), _ptr(ignore),
                                        _ptr(ign_offsets), _stream())
    _lib.check(rc, "cvm_prepare_objects")
    return objs, obj_offsets, ignore, ign_offsets


# ---- render ----------------------------------------------------------------------------------------------------------
def render_gt(layout: Layout, objs_dev, obj_offsets_dev, B, ignore_dev=None, ign_offsets_dev=None, out=None, extra=None,
              extra_off=0):
    """y_true [B,H,W,Ct] from device object records (see pack in models/centernet/processor.py).
    extra: optional float32 CUDA [n_obj, n] of per-object targets (l_shape / 3d_info) for channels [extra_off, extra_off+n)."""
    _need_cuda(objs_dev, "objs")
    _need_cuda(obj_offsets_dev, "obj_offsets", torch.int32)
    if out is None:
        out = torch.empty((B, layout.H, layout.W, layout.Ct), dtype=torch.float32, device=objs_dev.device)
    _need_cuda(out, "out", torch.float32)
    if tuple(out.shape) != (B, layout.H, layout.W, layout.Ct) or not out.is_contiguous():
        raise _lib.CvmError("out must be a contiguous [B,H,W,Ct] tensor")
    s = layout.c_struct()
    if extra is not None and extra.numel() > 0:
        _need_cuda(extra, "extra", torch.float32)
        if extra.dim() != 2 or not extra.is_contiguous():
            raise _lib.CvmError("extra must be a contiguous [n_obj, n] float32 tensor")
        rc = _lib.lib().cvm_render_gt_extra(C.byref(s), _ptr(objs_dev), _ptr(obj_offsets_dev), _ptr(ignore_dev),
                                            _ptr(ign_offsets_dev), B, _ptr(extra), int(extra.shape[1]), int(extra_off),
                                            int(extra.shape[1]), _ptr(out), _stream())
        _lib.check(rc, "cvm_render_gt_extra")
        return out
    rc = _lib.lib().cvm_render_gt(C.byref(s), _ptr(objs_dev), _ptr(obj_offsets_dev), _ptr(ignore_dev),
                                  _ptr(ign_offsets_dev), B, _ptr(out), _stream())
    _lib.check(rc, "cvm_render_gt")
    return out


def render_prev_heatmap(layout: Layout, objs_dev, obj_offsets_dev, B, out=None):
    _need_cuda(objs_dev, "objs")
    _need_cuda(obj_offsets_dev, "obj_offsets", torch.int32)
    if out is None:
        out = torch.empty((B, layout.H, layout.W, 1), dtype=torch.float32, device=objs_dev.device)
    _need_cuda(out, "out", torch.float32)
    if out.numel() != B * layout.H * layout.W or not out.is_contiguous():
        raise _lib.CvmError("out must be a contiguous [B,H,W,1] tensor")
    s = layout.c_struct()
    rc = _lib.lib().cvm_render_prev_hm(C.byref(s), _ptr(objs_dev), _ptr(obj_offsets_dev), B, _ptr(out), _stream())
    _lib.check(rc, "cvm_render_prev_hm")
    return out


def fill_heatmap_inplace(objs_dev, n_obj, heat, weights, H, W, R, alpha):
    """max/min-combine records into existing planes: heat [H,W,C] (channel 0 of the given view) and weights [H,W] or None."""
    _need_cuda(heat, "heat", torch.float32)
    _need_cuda(objs_dev, "objs")
    H, W = int(H), int(W)
    if heat.dim() == 3:       # [H,W,C] view: channel 0 of the given view, rows must be dense in pixels
        hs = int(heat.stride(-2))
        if tuple(heat.shape[:2]) != (H, W) or heat.stride(0) != W * hs or heat.stride(-1) != 1:
            raise _lib.CvmError("heat must be a [H,W,C] view with stride(0) == W*stride(1) and dense channels")
    elif heat.dim() == 2:
        hs = 1
        if tuple(heat.shape) != (H, W) or not heat.is_contiguous():
            raise _lib.CvmError("heat must be a contiguous [H,W] plane")
    else:
        raise _lib.CvmError("heat must be [H,W] or [H,W,C]")
    if weights is not None:
        _need_cuda(weights, "weights", torch.float32)
        if tuple(weights.shape) != (H, W) or not weights.is_contiguous():
            raise _lib.CvmError("weights must be a contiguous float32 [H,W] plane")
    if objs_dev.numel() < int(n_obj) * OBJ_DTYPE.itemsize:
        raise _lib.CvmError("objs holds fewer than n_obj records")
    rc = _lib.lib().cvm_fill_heatmap_inplace(_ptr(objs_dev), int(n_obj), _ptr(heat), hs, _ptr(weights), int(H), int(W),
                                             float(R), float(alpha), _stream())
    _lib.check(rc, "cvm_fill_heatmap_inplace")


# ---- loss ------------------------------------------------------------------------------------------------------------
def _n_pixels(t):
    n = 1
    for d in t.shape[:-1]:
        n *= int(d)
    return n


def loss_partials(layout: Layout, y_true, y_pred, use_weights=True, out=None):
    """One streaming pass -> fp64[16] device vector [P, N, n_pos, n_obj, field sums...] (sum it across GPUs, then finalize)."""
    st_t = _pixel_strided(y_true, "y_true")
    st_p = _pixel_strided(y_pred, "y_pred")
    n = _n_pixels(y_true)
    if _n_pixels(y_pred) != n:
        raise _lib.CvmError("y_true and y_pred must cover the same pixels")
    if out is None:
        out = torch.empty(_lib.CVM_NPART, dtype=torch.float64, device=y_true.device)
    s = layout.c_struct()
    nbytes = _lib.lib().cvm_loss_workspace_bytes(C.byref(s), n)
    ws = _workspace(y_true.device, nbytes, "loss")
    rc = _lib.lib().cvm_loss_fwd(C.byref(s), _ptr(y_true), st_t, _ptr(y_pred), st_p, n, int(bool(use_weights)),
                                 _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "cvm_loss_fwd")
    return out


def loss_total(layout: Layout, y_true, y_pred, use_weights=True, partials=None, out=None):
    """Single-device loss in ONE launch (cvm_loss_fwd_total): returns (out, partials) with out = float32 device vector
    [total, focal, field_0, ...] and partials the fp64[16] vector (needed by loss_backward)."""
    st_t = _pixel_strided(y_true, "y_true")
    st_p = _pixel_strided(y_pred, "y_pred")
    n = _n_pixels(y_true)
    if _n_pixels(y_pred) != n:
        raise _lib.CvmError("y_true and y_pred must cover the same pixels")
    if partials is None:
        partials = torch.empty(_lib.CVM_NPART, dtype=torch.float64, device=y_true.device)
    if out is None:
        out = torch.zeros(2 + _lib.CVM_MAX_FIELDS, dtype=torch.float32, device=y_true.device)
    s = layout.c_struct()
    nbytes = _lib.lib().cvm_loss_workspace_bytes(C.byref(s), n)
    ws = _workspace(y_true.device, nbytes, "loss")
    rc = _lib.lib().cvm_loss_fwd_total(C.byref(s), _ptr(y_true), st_t, _ptr(y_pred), st_p, n, int(bool(use_weights)),
                                       _ptr(partials), _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "cvm_loss_fwd_total")
    return out, partials


def loss_finalize(layout: Layout, partials, out=None):
    """-> float32 device vector [total, focal, field_0, ...] (fields normalised, post-transformed, unweighted)."""
    _need_cuda(partials, "partials", torch.float64)
    if out is None:
        out = torch.zeros(2 + _lib.CVM_MAX_FIELDS, dtype=torch.float32, device=partials.device)
    s = layout.c_struct()
    rc = _lib.lib().cvm_loss_finalize(C.byref(s), _ptr(partials), _ptr(out), _stream())
    _lib.check(rc, "cvm_loss_finalize")
    return out


def loss_finalize_gathered(layout: Layout, gathered, partials=None, out=None, finalize=True):
    """gathered [n_ranks,16] fp64 (every rank's partials) -> (out, partials): summed in rank order (bit-reproducible), finalised
    (finalize=False: only the sum; out is None then)."""
    _need_cuda(gathered, "gathered", torch.float64)
    if gathered.dim() != 2 or gathered.shape[1] != _lib.CVM_NPART or not gathered.is_contiguous():
        raise _lib.CvmError("gathered must be a contiguous [n_ranks, 16] float64 tensor")
    if partials is None:
        partials = torch.empty(_lib.CVM_NPART, dtype=torch.float64, device=gathered.device)
    if out is None and finalize:
        out = torch.zeros(2 + _lib.CVM_MAX_FIELDS, dtype=torch.float32, device=gathered.device)
    s = layout.c_struct()
    rc = _lib.lib().cvm_loss_finalize_gathered(C.byref(s), _ptr(gathered), int(gathered.shape[0]), _ptr(partials), _ptr(out), _stream())
    _lib.check(rc, "cvm_loss_finalize_gathered")
    return out, partials


def set_decode_spare_sms(n):
    """SMs the decode scan leaves free (1 for data-parallel callers whose collective should run beside the decode)."""
    _lib.check(_lib.lib().cvm_decode_set_spare_sms(int(n)), "cvm_decode_set_spare_sms")


def decode_fallback_count():
    """Images whose predicted decode threshold did not hold and that were recomputed by the slow exact path (monitoring)."""
    return int(_lib.lib().cvm_decode_fallback_count())


def loss_backward(layout: Layout, y_true, y_pred, partials, upstream=None, out=None, generic=False):
    """grad wrt y_pred[..., :Cp] as a contiguous [n_pixels, Cp] tensor; `partials` must be the globally reduced vector.
    generic=True: the layout-generic kernel only (cvm_loss_bwd_generic; tests compare it with the compile-time layouts)."""
    st_t = _pixel_strided(y_true, "y_true")
    st_p = _pixel_strided(y_pred, "y_pred")
    n = _n_pixels(y_true)
    if out is None:
        out = torch.empty(tuple(y_pred.shape[:-1]) + (layout.Cp,), dtype=torch.float32, device=y_pred.device)
    if upstream is not None:
        _need_cuda(upstream, "upstream", torch.float32)
    s = layout.c_struct()
    fn = _lib.lib().cvm_loss_bwd_generic if generic else _lib.lib().cvm_loss_bwd
    rc = fn(C.byref(s), _ptr(y_true), st_t, _ptr(y_pred), st_p, n, _ptr(partials), _ptr(upstream), _ptr(out), _stream())
    _lib.check(rc, "cvm_loss_bwd")
    return out


# ---- decode ----------------------------------------------------------------------------------------------------------
def decode_topk(layout: Layout, y_pred, K=100, rois_dev=None, want_track=None, semseg=None):
    """Canonical CenterNet decode -> dict(scores[B,K], cls[B,K] i32, flat[B,K] i64, centers[B,K,2], boxes[B,K,4], track[B,K,2]).
    semseg=(off, n_cls): also take the per-pixel argmax of channels [off, off+n_cls) of the same tensor in the same pass
    (multitask head) -> out["semseg_ids"] uint8 [B,H,W]."""
    st = _pixel_strided(y_pred, "y_pred")
    if y_pred.dim() != 4 or y_pred.shape[1] != layout.H or y_pred.shape[2] != layout.W:
        raise _lib.CvmError("y_pred must be [B,H,W,C] with the layout's H,W")
    B = int(y_pred.shape[0])
    dev = y_pred.device
    if want_track is None:
        want_track = layout.off_track >= 0
    out = dict(scores=torch.empty((B, K), dtype=torch.float32, device=dev),
               cls=torch.empty((B, K), dtype=torch.int32, device=dev),
               flat=torch.empty((B, K), dtype=torch.int64, device=dev),
               centers=torch.empty((B, K, 2), dtype=torch.float32, device=dev),
               boxes=torch.empty((B, K, 4), dtype=torch.float32, device=dev),
               track=torch.empty((B, K, 2), dtype=torch.float32, device=dev) if want_track else None)
    s = layout.c_struct()
    nbytes = _lib.lib().cvm_decode_topk_workspace_bytes(C.byref(s), st, B, K)
    if nbytes == 0 and B > 0:
        raise _lib.CvmError("cvm_decode_topk: unsupported shape: " + _lib.lib().cvm_last_error().decode())
    ws = _workspace(dev, nbytes, "decode")
    if semseg is None:
        rc = _lib.lib().cvm_decode_topk(C.byref(s), _ptr(y_pred), st, B, K, _ptr(rois_dev), _ptr(out["scores"]),
                                        _ptr(out["cls"]), _ptr(out["flat"]), _ptr(out["centers"]), _ptr(out["boxes"]),
                                        _ptr(out["track"]), _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "cvm_decode_topk")
    else:
        off, n_cls = int(semseg[0]), int(semseg[1])
        out["semseg_ids"] = torch.empty((B, layout.H, layout.W), dtype=torch.uint8, device=dev)
        rc = _lib.lib().cvm_decode_topk_semseg(C.byref(s), _ptr(y_pred), st, B, K, _ptr(rois_dev), _ptr(out["scores"]),
                                               _ptr(out["cls"]), _ptr(out["flat"]), _ptr(out["centers"]), _ptr(out["boxes"]),
                                               _ptr(out["track"]), off, n_cls, _ptr(out["semseg_ids"]), _ptr(ws), ws.numel(),
                                               _stream())
        _lib.check(rc, "cvm_decode_topk_semseg")
    return out


def decode_window9(layout: Layout, y_pred, min_conf=0.25, rois_dev=None, max_out=256, window=9):
    """The reference's process_2d_output on a batch -> dict(counts[B], cls, pix, scores, centers, boxes) in scan order."""
    st = _pixel_strided(y_pred, "y_pred")
    B = int(y_pred.shape[0])
    dev = y_pred.device
    out = dict(counts=torch.zeros(B, dtype=torch.int32, device=dev),
               cls=torch.zeros((B, max_out), dtype=torch.int32, device=dev),
               pix=torch.zeros((B, max_out), dtype=torch.int32, device=dev),
               scores=torch.zeros((B, max_out), dtype=torch.float32, device=dev),
               centers=torch.zeros((B, max_out, 2), dtype=torch.float32, device=dev),
               boxes=torch.zeros((B, max_out, 4), dtype=torch.float32, device=dev))
    s = layout.c_struct()
    ws = _workspace(dev, _lib.lib().cvm_decode_window9_workspace_bytes(C.byref(s), B), "window9")
    rc = _lib.lib().cvm_decode_window9(C.byref(s), _ptr(y_pred), st, B, window, float(min_conf), _ptr(rois_dev), max_out,
                                       _ptr(out["counts"]), _ptr(out["cls"]), _ptr(out["pix"]), _ptr(out["scores"]),
                                       _ptr(out["centers"]), _ptr(out["boxes"]), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "cvm_decode_window9")
    return out


# ---- CenterTrack association -----------------------------------------------------------------------------------------
def track_associate(det, prev_centers, prev_sizes, prev_cls, prev_count=None, min_score=0.0):
    """Greedy CenterTrack association (cvm_track_associate).  det: the dict of decode_topk (centers, track, boxes, scores,
    cls); prev_*: [B,M,2] centres, [B,M,2] (w,h), [B,M] classes of the previous frame's tracks, prev_count [B] valid rows.
    Returns match [B,K] int32: index of the matched previous track or -1."""
    if det.get("track") is None:
        raise _lib.CvmError("track_associate needs the track output of decode_topk (a layout with track_offset)")
    B, K = (int(v) for v in det["scores"].shape)
    M = int(prev_centers.shape[1]) if prev_centers.dim() == 3 else 0
    for name, t, dt in (("centers", det["centers"], torch.float32), ("track", det["track"], torch.float32),
                        ("boxes", det["boxes"], torch.float32), ("scores", det["scores"], torch.float32),
                        ("cls", det["cls"], torch.int32), ("prev_centers", prev_centers, torch.float32),
                        ("prev_sizes", prev_sizes, torch.float32), ("prev_cls", prev_cls, torch.int32)):
        _need_cuda(t, name, dt)
        if not t.is_contiguous():
            raise _lib.CvmError(f"{name} must be contiguous")
    if prev_count is not None:
        _need_cuda(prev_count, "prev_count", torch.int32)
    if tuple(prev_centers.shape) != (B, M, 2) or tuple(prev_sizes.shape) != (B, M, 2) or tuple(prev_cls.shape) != (B, M):
        raise _lib.CvmError("previous-frame tensors must be [B,M,2], [B,M,2], [B,M]")
    match = torch.empty((B, K), dtype=torch.int32, device=det["scores"].device)
    rc = _lib.lib().cvm_track_associate(_ptr(det["centers"]), _ptr(det["track"]), _ptr(det["boxes"]), _ptr(det["scores"]),
                                        _ptr(det["cls"]), B, K, float(min_score), _ptr(prev_centers), _ptr(prev_sizes),
                                        _ptr(prev_cls), _ptr(prev_count), M, _ptr(match), _stream())
    _lib.check(rc, "cvm_track_associate")
    return match


# ---- semseg argmax ---------------------------------------------------------------------------------------------------
def semseg_argmax(x, off, n_cls, lut_bgr=None, threshold=None, use_weight=False, apply_softmax=True):
    """x [...,C] float32 CUDA. lut_bgr None -> uint8 class ids [...]; else uint8 BGR [...,3] with to_3channel semantics."""
    st = _pixel_strided(x, "x")
    n = _n_pixels(x)
    if lut_bgr is None:
        out = torch.empty(tuple(x.shape[:-1]), dtype=torch.uint8, device=x.device)
        mode = 0
    else:
        _need_cuda(lut_bgr, "lut_bgr", torch.uint8)
        out = torch.empty(tuple(x.shape[:-1]) + (3,), dtype=torch.uint8, device=x.device)
        mode = 1
    thr = float("nan") if threshold is None else float(threshold)
    rc = _lib.lib().cvm_semseg_argmax(_ptr(x), n, st, int(off), int(n_cls), mode, int(bool(apply_softmax)),
                                      int(bool(use_weight)), thr, _ptr(lut_bgr), _ptr(out), _stream())
    _lib.check(rc, "cvm_semseg_argmax")
    return out
