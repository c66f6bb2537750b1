"""CPU: the oracle restatements against the committed golden vectors (generated from the REAL reference functions by
tests/golden/make_golden.py) and, where /root/reference is mounted, against the live reference functions."""
import os

import numpy as np
import pytest

from oracle import decode_np, image_np, ref_import, render_np
from oracle.layout import make_layout


def _host_filter(raw_boxes, in_w, in_h, min_area):
    """clip_to_img + MIN_BOX_AREA filter, restated from processor.py:46-56,241-253."""
    keep, ign = [], []
    for i, (x, y, w, h) in enumerate(raw_boxes):
        mx, my = np.clip(x + w, 0, in_w), np.clip(y + h, 0, in_h)
        x, y = np.clip(x, 0, in_w), np.clip(y, 0, in_h)
        cb = [x, y, mx - x, my - y]
        if cb[2] * cb[3] > min_area:
            keep.append((i, cb))
        else:
            ign.append(cb)
    return keep, ign


@pytest.mark.parametrize("name", ["render_process_a", "render_process_b", "render_process_empty"])
def test_render_vs_real_process(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    in_w, in_h, R = int(g["in_w"]), int(g["in_h"]), int(g["R"])
    L = make_layout(in_h // R, in_w // R, 6, "R", R=R, alpha=float(g["alpha"]))
    keep, ign = _host_filter(g["raw_boxes"], in_w, in_h, float(g["min_box_area"]))
    y = render_np.render_image(L, [cb for _, cb in keep], [g["cls"][i] for i, _ in keep], ign)
    assert y.shape == g["y_true"].shape
    assert np.array_equal(y, g["y_true"])


def test_fill_vs_real_fill_heatmap(golden_dir):
    g = np.load(os.path.join(golden_dir, "fill_cases.npz"))
    H, W = int(g["H"]), int(g["W"])
    heat = np.zeros((H, W), np.float32)
    wts = np.ones((H, W), np.float32)
    for cx, cy, w, h, peak in g["recs"]:
        render_np.fill(heat, wts, 0.9, 2, int(cx), int(cy), w, h, W, H, peak)
    assert np.array_equal(heat, g["heat"][..., 0])
    assert np.array_equal(wts, g["wts"])
    assert wts.min() < 1.0 and heat.max() == 1.0


def test_decode_window9_vs_real(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode_r.npz"))
    m = g["mask"]
    L = make_layout(m.shape[0], m.shape[1], int(g["nb_classes"]), "R")
    objs = decode_np.decode_window9(L, m, tuple(g["roi"]), float(g["min_conf"]))
    assert len(objs) == len(g["cls"]) > 5
    assert [o["cls_idx"] for o in objs] == list(g["cls"])
    assert np.array_equal(np.array([o["center"] for o in objs], np.float32), g["center"])
    assert np.array_equal(np.array([o["fullbox"] for o in objs], np.float32), g["fullbox"])
    # fixture semantics: plateau keeps the first pixel only, dead border, strict threshold, class tie -> first
    pos = {(o["y"], o["x"]) for o in objs}
    assert (10, 20) in pos and (10, 21) not in pos and (3, 30) not in pos and (25, 40) not in pos
    assert [o["cls_idx"] for o in objs if (o["y"], o["x"]) == (30, 50)] == [0]


def test_to_3channel_vs_real(golden_dir):
    g = np.load(os.path.join(golden_dir, "to3.npz"))
    cols = [tuple(int(v) for v in c) for c in g["colours"]]
    for thr in (None, 0.3):
        for uw in (False, True):
            for sm in (True, False):
                key = f"out_thr{'N' if thr is None else '1'}_w{int(uw)}_s{int(sm)}"
                out = image_np.to_3channel(g["raw"], cols, thr, uw, sm)
                assert np.array_equal(out, g[key]), key


@pytest.mark.skipif(not ref_import.available(), reason="reference not mounted")
def test_live_reference_random():
    """Fresh seeds against the live functions (build container only)."""
    r = ref_import.load()
    rng = np.random.default_rng(77)
    H, W = 64, 96
    heat = np.zeros((H, W, 3), np.float32)
    wts = np.ones((H, W), np.float32)
    heat2, wts2 = heat.copy(), wts.copy()
    for _ in range(30):
        cx, cy = int(rng.integers(0, W)), int(rng.integers(0, H))
        w, h = float(np.exp(rng.uniform(0.3, 5.2))), float(np.exp(rng.uniform(0.3, 4.6)))
        r["fill_heatmap"](heat, 0.9, 2, wts, cx, cy, w, h, W, H)
        render_np.fill(heat2[:, :, 0], wts2, 0.9, 2, cx, cy, w, h, W, H)
    assert np.array_equal(heat, heat2) and np.array_equal(wts, wts2)

    P = r["CenternetParams"](4)
    L = make_layout(40, 56, 4, "R")
    m = rng.normal(0, 1, (40, 56, L.Cp)).astype(np.float32)
    m[..., 0] = 1 / (1 + np.exp(-rng.normal(-1.5, 1.5, (40, 56))))
    roi = r["Roi"]()
    roi.scale, roi.offset_left, roi.offset_top = 0.5, 3, -7
    ref = r["process_2d_output"](m, roi, P, 0.3)
    mine = decode_np.decode_window9(L, m, (0.5, 3, -7), 0.3)
    assert len(ref) == len(mine) > 0
    for a, b in zip(ref, mine):
        assert int(a["cls_idx"]) == b["cls_idx"]
        assert np.array_equal(np.float32(a["center"]), np.float32(b["center"]))
        assert np.array_equal(np.float32(a["fullbox"]), np.float32(b["fullbox"]))
