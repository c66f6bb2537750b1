"""TEST INFRASTRUCTURE ONLY — CPU restatement of the semseg colour/argmax step.

Follows `to_3channel`, /root/reference/common/utils/image.py:72-100, vectorised over pixels.
Validated against the REAL numba function in tests/test_oracle_golden.py::test_live_reference_random and the fixtures
tests/golden/to3_*.npz. Behaviour pinned there: ties -> first index; an all-equal
row with apply_softmax=True becomes 0/0 = NaN -> argmax 0; NaN fails `> threshold`; the uint8 cast
truncates. Unlike the reference this restatement does NOT mutate its input (image.py:82 does, quirk C.9).
"""
import numpy as np


def class_ids(raw, n_cls):
    """argmax over the first n_cls channels (first max wins, image.py:89) -> uint8 [H,W]."""
    a = np.asarray(raw)[..., :n_cls]
    return np.argmax(a, axis=-1).astype(np.uint8)


def to_3channel(raw, colours, threshold=None, use_weight=False, apply_softmax=True):
    """raw [H,W,>=n_cls] float; colours: sequence of n_cls (B,G,R) tuples. Returns uint8 [H,W,3]."""
    n_cls = len(colours)
    raw = np.asarray(raw)
    H, W = raw.shape[:2]
    arr = raw.reshape(-1, raw.shape[2])[:, :n_cls].astype(raw.dtype, copy=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        if apply_softmax:
            arr = arr - arr.min(axis=1, keepdims=True)                      # :82
            arr = arr / arr.sum(axis=1, keepdims=True)                      # :83
        if n_cls == 1:
            idx = np.zeros(arr.shape[0], dtype=np.int64)                    # :85-87
            score = arr[:, 0]
        else:
            idx = np.argmax(arr, axis=1)                                    # :89
            picked = arr[np.arange(arr.shape[0]), idx]
            # numba evaluates min(1.0, max(0.0, nan)) to 0.0 (nan compares false), checked against the real function
            score = np.where(np.isnan(picked), 0.0, np.minimum(1.0, np.maximum(0.0, picked)))
    lut = np.asarray(colours, dtype=np.float64).reshape(n_cls, 3)
    if threshold is None:
        keep = np.ones(arr.shape[0], dtype=bool)
    else:
        with np.errstate(invalid="ignore"):
            keep = score.astype(np.float64) > float(threshold)              # :92 (numba promotes the f32 score to f64)
    s = score.astype(np.float64) if use_weight else np.ones(arr.shape[0])   # :94
    col = lut[idx] * s[:, None]                                             # :96
    with np.errstate(invalid="ignore"):
        col8 = np.where(np.isnan(col), 0, col).astype(np.int64).astype(np.uint8)   # :99 truncating cast
    out = np.where(keep[:, None], col8, 0).astype(np.uint8)
    return out.reshape(H, W, 3)
