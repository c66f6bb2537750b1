"""Seeded synthetic inputs of SURVEY.md section 8(d) (NumPy, host side) shared by the parity tests."""
import numpy as np


def objects_for_image(rng, H, W, R, n_obj, K):
    """Boxes in INPUT px (fp64, already inside the image), class ids."""
    in_w, in_h = W * R, H * R
    boxes, cls = [], []
    for _ in range(n_obj):
        w = float(np.exp(rng.uniform(np.log(4), np.log(min(160, in_w / 2)))))
        h = float(np.exp(rng.uniform(np.log(4), np.log(min(96, in_h / 2)))))
        cx, cy = float(rng.uniform(0, in_w)), float(rng.uniform(0, in_h))
        x0, y0 = max(0.0, cx - w / 2), max(0.0, cy - h / 2)
        x1, y1 = min(float(in_w), cx + w / 2), min(float(in_h), cy + h / 2)
        boxes.append([x0, y0, x1 - x0, y1 - y0])
        cls.append(int(rng.integers(0, K)))
    return np.asarray(boxes, np.float64).reshape(-1, 4), np.asarray(cls, np.int32)


def ignore_for_image(rng, H, W, n):
    out = []
    for _ in range(n):
        out.append([float(rng.uniform(0, W - 4)), float(rng.uniform(0, H - 4)), float(rng.uniform(1, 12)), float(rng.uniform(1, 8))])
    return np.asarray(out, np.float64).reshape(-1, 4)


def prediction_for_image(rng, L, boxes, cls, ties=True):
    """y_pred [H,W,Cp] f32: sigmoid noise heatmaps with object centres raised, regression heads filled."""
    H, W, hm = L.H, L.W, L.hm
    yp = np.zeros((H, W, L.Cp), np.float32)
    yp[..., :hm] = (1.0 / (1.0 + np.exp(-rng.normal(-4.0, 1.5, (H, W, hm))))).astype(np.float32)
    if L.off_class >= 0:
        yp[..., L.off_class:L.off_class + L.nb_classes] = rng.normal(0, 1, (H, W, L.nb_classes)).astype(np.float32)
    if L.off_roff >= 0:
        yp[..., L.off_roff:L.off_roff + 2] = rng.uniform(0, 1, (H, W, 2)).astype(np.float32)
    if L.off_box >= 0:
        yp[..., L.off_box:L.off_box + 2] = rng.uniform(4, 120, (H, W, 2)).astype(np.float32)
    if L.off_track >= 0:
        yp[..., L.off_track:L.off_track + 2] = rng.normal(0, 8, (H, W, 2)).astype(np.float32)
    for f in L.fields:
        if f.name in ("l_shape", "radial_dist", "orientation", "obj_dims"):
            yp[..., f.off:f.off + f.size] = rng.normal(0, 2, (H, W, f.size)).astype(np.float32)
    for b, c in zip(boxes, cls):
        cx = max(0, min(W - 1, int((b[0] + b[2] / 2) / L.R)))
        cy = max(0, min(H - 1, int((b[1] + b[3] / 2) / L.R)))
        ch = int(c) if hm > 1 else 0
        yp[cy, cx, ch] = np.float32(rng.uniform(0.3, 0.99))
        if L.off_box >= 0:
            yp[cy, cx, L.off_box:L.off_box + 2] = (np.asarray(b[2:4]) * np.exp(rng.normal(0, 0.1, 2))).astype(np.float32)
    if ties and H >= 8 and W >= 8:
        # duplicate one peak value into 3 other positions / classes and add a 2x2 plateau
        v = np.float32(0.93)
        for _ in range(4):
            yp[int(rng.integers(0, H)), int(rng.integers(0, W)), int(rng.integers(0, hm))] = v
        py, px, pc = int(rng.integers(0, H - 1)), int(rng.integers(0, W - 1)), int(rng.integers(0, hm))
        yp[py:py + 2, px:px + 2, pc] = np.float32(0.97)
    return yp


def make_batch(L, config, B, n_obj=None, n_ignore=2, track=False, ties=True):
    """Returns dict(boxes[list], cls[list], ignore[list], track[list or None], y_pred [B,H,W,Cp])."""
    out = dict(boxes=[], cls=[], ignore=[], track=[] if track else None, y_pred=[])
    for i in range(B):
        rng = np.random.default_rng(1234 + 1000 * config + i)
        n = int(rng.integers(0, 65)) if n_obj is None else n_obj
        boxes, cls = objects_for_image(rng, L.H, L.W, L.R, n, L.nb_classes)
        out["boxes"].append(boxes)
        out["cls"].append(cls)
        out["ignore"].append(ignore_for_image(rng, L.H, L.W, n_ignore))
        if track:
            out["track"].append(rng.normal(0, 8, (len(boxes), 2)).astype(np.float32))
        out["y_pred"].append(prediction_for_image(rng, L, boxes, cls, ties))
    out["y_pred"] = np.stack(out["y_pred"])
    return out
