"""Decode of the CenterNet output.

`process_2d_output` keeps the reference signature and result format (models/centernet/post_processing.py:6-66: list of
dicts with cls_idx / center / fullbox, scan order, 9x9 first-argmax window, strict threshold) but runs in
cvm_decode_window9.  `decode_topk` is the canonical batched decode (3x3 NMS + per-image top-K) in cvm_decode_topk.
"""
import numpy as np
import torch

from cvmhot import ops
from cvmhot.layout import layout_from_params


def _roi_tuple(roi):
    return (float(roi.scale), float(roi.offset_left), float(roi.offset_top))


def process_2d_output(output_mask, roi, params, min_conf_value=0.25, max_objects=512):
    """output_mask: [H,W,C] numpy array or CUDA tensor of ONE image (no batch dim, like the reference).
    `max_objects` only sizes the first attempt; if the map holds more peaks the decode runs again with room for all of
    them (the reference returns every peak).  The map is read as fp32 (what the network emits); box math is fp32
    op-for-op, which equals the reference under NumPy >= 2 scalar promotion (python float * np.float32 -> float32;
    legacy NumPy < 2 promoted to float64 there, see tests/golden/README)."""
    is_np = not isinstance(output_mask, torch.Tensor)
    x = torch.from_numpy(np.ascontiguousarray(output_mask, dtype=np.float32)).cuda() if is_np else output_mask
    H, W = int(x.shape[0]), int(x.shape[1])
    L = layout_from_params(params, H=H, W=W)
    rois_dev = ops.make_rois([_roi_tuple(roi)], x.device)
    out = ops.decode_window9(L, x.unsqueeze(0), min_conf=min_conf_value, rois_dev=rois_dev, max_out=max_objects)
    n = int(out["counts"][0].item())
    if n > max_objects:       # the reference returns EVERY peak: run again with room for all of them
        out = ops.decode_window9(L, x.unsqueeze(0), min_conf=min_conf_value, rois_dev=rois_dev, max_out=n)
    cls = out["cls"][0, :n].cpu().numpy()
    centers = out["centers"][0, :n].cpu().numpy()
    boxes = out["boxes"][0, :n].cpu().numpy()
    pix = out["pix"][0, :n].cpu().numpy()
    objects = []
    fields = params.REGRESSION_FIELDS
    host = output_mask if is_np else None
    for i in range(n):
        obj = {"cls_idx": int(cls[i]) if fields["class"].active else 0, "center": [centers[i, 0], centers[i, 1]]}
        if fields["fullbox"].active:
            obj["fullbox"] = [boxes[i, 0], boxes[i, 1], boxes[i, 2], boxes[i, 3]]
        if fields["l_shape"].active:      # reference :54-59, cheap host glue on the few detected pixels
            if host is None:
                host = x.cpu().numpy()
            y_, x_ = divmod(int(pix[i]), W)
            ls = host[y_, x_, params.start_idx("l_shape"):params.end_idx("l_shape")]
            inv = 1.0 / roi.scale
            c = np.asarray(obj["center"])
            obj["bottom_left"] = c + ls[0:2] * inv
            obj["bottom_right"] = c + ls[2:4] * inv
            obj["bottom_center"] = c + ls[4:6] * inv
            obj["center_height"] = ls[6] * inv
        objects.append(obj)
    return objects


def decode_topk(y_pred, params, K=100, rois=None):
    """y_pred [B,H,W,C] CUDA tensor -> dict of CUDA tensors (scores, cls, flat, centers, boxes, track).
    rois: optional list of Roi (one per image)."""
    L = layout_from_params(params, H=int(y_pred.shape[1]), W=int(y_pred.shape[2]))
    rois_dev = ops.make_rois([_roi_tuple(r) for r in rois], y_pred.device) if rois is not None else None
    return ops.decode_topk(L, y_pred, K=K, rois_dev=rois_dev)
