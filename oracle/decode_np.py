"""TEST INFRASTRUCTURE ONLY — CPU restatement of the output decode (never on the product path).

Profile R (the reference as shipped):
  process_2d_output     /root/reference/models/centernet/post_processing.py:6-66
  convert_back_to_roi   /root/reference/common/utils/image.py:22-28
  `decode_window9` is a vectorised restatement and is validated against the REAL function
  (tests/test_oracle_golden.py::test_live_reference_random, tests/golden/decode_r.npz).

Profile N (what BASELINE.json's north_star specifies; SURVEY.md App. A.3.2 — there is no such code
in the reference, so parity for it is defined by this file and is otherwise UNPINNED):
  S = v where v equals the max of its in-bounds 3x3 neighbourhood (same channel) and v > 0, else 0
  candidates ordered by (S desc, NHWC flat index asc) == tf.nn.top_k tie rule; first K taken;
  regression heads gathered at the peak pixel; boxes assembled with the reference's fullbox math
  (post_processing.py:44-52) in fp32.
"""
import numpy as np

from .layout import Layout


def _f32(x):
    return np.float32(x)


def box_math(x, y, dx, dy, w, h, R, scale, off_left, off_top):
    """post_processing.py:44-52 + image.py:22-28 in fp32 (NumPy weak-scalar promotion keeps f32).

    centre = (1/scale) * ((x + dx) * R) - offset ; w,h *= (1/scale) ; box = [cx - w/2, cy - h/2, w, h]
    """
    inv = _f32(1.0 / scale)
    px = (_f32(x) + _f32(dx)) * _f32(R)
    py = (_f32(y) + _f32(dy)) * _f32(R)
    cx = inv * px - _f32(off_left)
    cy = inv * py - _f32(off_top)
    bw = _f32(w) * inv
    bh = _f32(h) * inv
    return cx, cy, np.array([cx - bw / _f32(2.0), cy - bh / _f32(2.0), bw, bh], dtype=np.float32)


def nms3_scores(hmap):
    """hmap [H,W,hm] f32 -> S [H,W,hm] f32: 3x3 max-pool keep (SAME, -inf padding), plateaus all kept."""
    H, W, C = hmap.shape
    pad = np.full((H + 2, W + 2, C), -np.inf, dtype=np.float32)
    pad[1:-1, 1:-1] = hmap
    mx = hmap.copy()
    for dy in range(3):
        for dx in range(3):
            mx = np.maximum(mx, pad[dy:dy + H, dx:dx + W])
    return np.where((hmap == mx) & (hmap > 0), hmap, np.float32(0)).astype(np.float32)


def decode_topk_image(L: Layout, y_pred, K=100, roi=(1.0, 0, 0)):
    """One image. y_pred [H,W,Cp] f32. roi = (scale, offset_left, offset_top).

    returns dict(scores[K] f32, cls[K] i32, flat[K] i64, centers[K,2] f32, boxes[K,4] f32, track[K,2] f32)
    `track` = tracking offset in roi coordinates (tx/scale, ty/scale) when the layout has a track field:
    centers + track is the predicted centre in the previous frame (what oracle/track_np.py consumes).
    """
    H, W, hm = L.H, L.W, L.hm
    y_pred = np.asarray(y_pred, dtype=np.float32)
    S = nms3_scores(y_pred[..., :hm]).reshape(-1)
    n = S.size
    K = min(K, n)
    # (S desc, flat asc): stable argsort on -S keeps ascending flat index inside ties
    order = np.argsort(-S.astype(np.float64), kind="stable")[:K]
    out = dict(scores=S[order].astype(np.float32), cls=(order % hm).astype(np.int32),
               flat=order.astype(np.int64), centers=np.zeros((K, 2), np.float32),
               boxes=np.zeros((K, 4), np.float32), track=np.zeros((K, 2), np.float32))
    scale, off_left, off_top = roi
    inv = _f32(1.0 / scale)
    for i, f in enumerate(order):
        x = (f // hm) % W
        y = f // (hm * W)
        px = y_pred[y, x]
        dx, dy = (px[L.off_roff], px[L.off_roff + 1]) if L.off_roff >= 0 else (0.0, 0.0)
        w, h = (px[L.off_box], px[L.off_box + 1]) if L.off_box >= 0 else (0.0, 0.0)
        if hm == 1 and L.off_class >= 0:        # Profile R: class = first argmax of the class logits (post_processing.py:39-41)
            out["cls"][i] = int(np.argmax(px[L.off_class:L.off_class + L.nb_classes]))
        cx, cy, box = box_math(x, y, dx, dy, w, h, L.R, scale, off_left, off_top)
        out["centers"][i] = (cx, cy)
        out["boxes"][i] = box
        if L.off_track >= 0:
            out["track"][i] = (px[L.off_track] * inv, px[L.off_track + 1] * inv)
    return out


def decode_topk(L, y_pred, K=100, rois=None):
    B = y_pred.shape[0]
    outs = [decode_topk_image(L, y_pred[b], K, (1.0, 0, 0) if rois is None else rois[b]) for b in range(B)]
    return {k: np.stack([o[k] for o in outs]) for k in outs[0]}


def decode_window9(L: Layout, output_mask, roi=(1.0, 0, 0), min_conf=0.25, win=9):
    """Profile R, post_processing.py:18-65, vectorised. output_mask [H,W,Cp] f32.

    A pixel (y,x) with win//2 <= y < H-win//2 (same for x) is an object iff np.argmax of the win x win window
    of channel 0 is the window centre (FIRST max in row-major order wins, :32) and value > min_conf (strict, :35).
    Returns a list of dicts in scan order with keys cls_idx, center, fullbox (+ score, y, x for testing).
    """
    m = np.asarray(output_mask)
    H, W = m.shape[:2]
    r = win // 2
    hm0 = m[:, :, 0]
    objs = []
    scale, off_left, off_top = roi
    for y in range(r, H - r):
        for x in range(r, W - r):
            v = hm0[y, x]
            if not (v > min_conf):
                continue
            wv = hm0[y - r:y + r + 1, x - r:x + r + 1].reshape(-1)
            c = r * win + r
            # first-argmax == centre: strictly greater than everything before, >= everything after
            if not (np.all(wv[:c] < v) and np.all(wv[c + 1:] <= v)):
                continue
            px = m[y, x]
            cls_idx = int(np.argmax(px[L.off_class:L.off_class + L.nb_classes])) if L.off_class >= 0 else 0   # :39-41
            dx, dy = (px[L.off_roff], px[L.off_roff + 1]) if L.off_roff >= 0 else (0.0, 0.0)
            w, h = (px[L.off_box], px[L.off_box + 1]) if L.off_box >= 0 else (0.0, 0.0)
            cx, cy, box = box_math(x, y, dx, dy, w, h, L.R, scale, off_left, off_top)                          # :43-52
            objs.append(dict(cls_idx=cls_idx, center=[cx, cy], fullbox=list(box), score=v, y=y, x=x))
    return objs
