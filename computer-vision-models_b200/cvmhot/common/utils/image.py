"""Roi bookkeeping and the semseg colour/argmax step (reference common/utils/image.py:9-28, 72-100).

`to_3channel` keeps the reference signature; the per-pixel work runs in cvm_semseg_argmax on the GPU.
"""
from dataclasses import dataclass

import numpy as np
import torch

from cvmhot import ops


@dataclass
class Roi:
    """Crop/resize description of the network input relative to the original image (offsets are pre-scaling)."""
    offset_top: int = 0
    offset_bottom: int = 0
    offset_left: int = 0
    offset_right: int = 0
    scale: float = 1.0


def convert_back_to_roi(roi: Roi, point):
    """Network-input coordinates -> original-image coordinates (reference image.py:22-28)."""
    inv = 1 / roi.scale
    return [inv * point[0] - roi.offset_left, inv * point[1] - roi.offset_top]


def _colours(cls_items):
    # accepts an OrderedDict name->(B,G,R), a list of (name, (B,G,R)) pairs (numba typed List in the reference) or colours
    if hasattr(cls_items, "values"):
        cols = list(cls_items.values())
    else:
        cols = [c[1] if (len(c) == 2 and not np.isscalar(c[1])) else c for c in cls_items]
    return np.asarray(cols, dtype=np.uint8).reshape(-1, 3)


def to_3channel(raw_mask_output, cls_items, threshold=None, use_weight=False, apply_softmax=True):
    """[H,W,>=n_cls] float -> uint8 BGR [H,W,3].  Accepts a numpy array (returned as numpy, like the reference) or a
    CUDA tensor (returned as a CUDA tensor, no host round trip).  Unlike the reference the input is never modified."""
    lut = _colours(cls_items)
    n_cls = lut.shape[0]
    is_np = not isinstance(raw_mask_output, torch.Tensor)
    x = torch.as_tensor(np.ascontiguousarray(raw_mask_output, dtype=np.float32)).cuda() if is_np else raw_mask_output
    out = ops.semseg_argmax(x, 0, n_cls, lut_bgr=torch.from_numpy(lut.reshape(-1)).to(x.device), threshold=threshold,
                            use_weight=use_weight, apply_softmax=apply_softmax)
    return out.cpu().numpy() if is_np else out


def class_ids(raw_mask_output, off, n_cls):
    """argmax only: uint8 class ids for channels [off, off+n_cls) of a CUDA tensor [...,C]."""
    return ops.semseg_argmax(raw_mask_output, off, n_cls)
