"""Generates the committed golden fixtures in this directory by running the REAL reference functions
(imported from /root/reference under module stubs, see oracle/ref_import.py). Run in the build container:

    python tests/golden/make_golden.py

The fixtures travel to the GPU box (where /root/reference does not exist); tests compare both the CPU
oracle and the CUDA path against them. Nothing here is imported by the product.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402

NAMES = ["car", "truck", "van", "motorbike", "cyclist", "ped"]   # data/label_spec.py:6-14 order


def gen_render_process(r, seed, n_obj, in_w, in_h):
    """Real ProcessImages.process (processor.py:215-339) on a blank PNG + synthetic object list."""
    import cv2
    P = r["CenternetParams"](6)
    P.INPUT_WIDTH, P.INPUT_HEIGHT = in_w, in_h
    proc = r["ProcessImages"](P)
    ok, buf = cv2.imencode(".png", np.zeros((in_h, in_w, 3), np.uint8))
    rng = np.random.default_rng(seed)
    raw_boxes, cls = [], []
    for i in range(n_obj):
        if i % 7 == 3:      # tiny boxes -> area <= MIN_BOX_AREA -> ignore areas (processor.py:243-253)
            w, h = float(rng.uniform(2, 3.8)), float(rng.uniform(2, 3.8))
            raw_boxes.append([float(rng.uniform(0, in_w / 2 - 5)), float(rng.uniform(0, in_h / 2 - 5)), w, h])
            cls.append(int(rng.integers(0, 6)))
            continue
        else:
            w = float(np.exp(rng.uniform(np.log(2), np.log(in_w / 2))))
            h = float(np.exp(rng.uniform(np.log(2), np.log(in_h / 1.5))))
        x, y = float(rng.uniform(-20, in_w)), float(rng.uniform(-15, in_h))
        raw_boxes.append([x, y, w, h])
        cls.append(int(rng.integers(0, 6)))
    if n_obj >= 2:          # two objects sharing one centre pixel (last writer wins, two class bits)
        raw_boxes[1] = [raw_boxes[0][0] + 0.25, raw_boxes[0][1] + 0.25, raw_boxes[0][2], raw_boxes[0][3]]
        cls[1] = (cls[0] + 1) % 6
    objs = [{"box2d": b, "obj_class": NAMES[c], "box3d_valid": False} for b, c in zip(raw_boxes, cls)]
    _, _, gt, _ = proc.process({"img": buf.tobytes(), "objects": objs}, None, None, {"epoch": 0})
    return dict(raw_boxes=np.array(raw_boxes, np.float64).reshape(-1, 4), cls=np.array(cls, np.int32),
                y_true=gt.astype(np.float32), in_w=in_w, in_h=in_h, min_box_area=P.MIN_BOX_AREA,
                R=P.R, alpha=P.VARIANCE_ALPHA)


def gen_fill_cases(r, seed):
    """Real fill_heatmap (processor.py:17-38): explicit centres (also outside the map) and peaks < 1."""
    H, W = 40, 56
    rng = np.random.default_rng(seed)
    heat = np.zeros((H, W, 1), np.float32)
    wts = np.ones((H, W), np.float32)
    recs = []
    for i in range(24):
        cx, cy = int(rng.integers(-6, W + 6)), int(rng.integers(-6, H + 6))
        w = float(np.exp(rng.uniform(np.log(1.2), np.log(90))))
        h = float(np.exp(rng.uniform(np.log(1.2), np.log(60))))
        peak = 1.0 if i % 3 == 0 else float(rng.uniform(0.04, 0.9))
        r["fill_heatmap"](heat, 0.9, 2, wts, cx, cy, w, h, W, H, peak)
        recs.append([cx, cy, w, h, peak])
    return dict(recs=np.array(recs, np.float64), heat=heat, wts=wts, H=H, W=W)


def gen_decode_r(r, seed):
    """Real process_2d_output (post_processing.py:6-66) on an fp32 map with plateaus and border peaks."""
    nb = 3
    P = r["CenternetParams"](nb)
    Cp = P.mask_channels()
    H, W = 40, 64
    rng = np.random.default_rng(seed)
    m = np.zeros((H, W, Cp), np.float32)
    m[..., 0] = (1.0 / (1.0 + np.exp(-rng.normal(-2.0, 1.5, (H, W))))).astype(np.float32)
    m[..., 1:] = rng.normal(0, 1, (H, W, Cp - 1)).astype(np.float32)
    m[..., P.start_idx("fullbox"):P.end_idx("fullbox")] = rng.uniform(4, 80, (H, W, 2)).astype(np.float32)
    m[..., P.start_idx("r_offset"):P.end_idx("r_offset")] = rng.uniform(0, 1, (H, W, 2)).astype(np.float32)
    m[10, 20, 0] = m[10, 21, 0] = 0.97          # 2-px plateau: only the first survives
    m[3, 30, 0] = 0.99                          # inside the 4-px dead border: ignored
    m[25, 40, 0] = 0.25                         # == threshold: rejected (strict >)
    m[24:27, 39:42, 0] = np.minimum(m[24:27, 39:42, 0], 0.2)
    m[25, 40, 0] = 0.25
    m[30, 50, 1:4] = [0.5, 0.5, 0.1]            # class tie -> first index
    m[30, 50, 0] = 0.95
    Roi = r["Roi"]
    roi = Roi()
    roi.scale, roi.offset_left, roi.offset_top = 0.25, -5, -3      # post_processing_test.py:47-51
    objs = r["process_2d_output"](m, roi, P, 0.25)
    return dict(mask=m, nb_classes=nb, roi=np.array([0.25, -5, -3], np.float64), min_conf=0.25,
                cls=np.array([int(o["cls_idx"]) for o in objs], np.int32),
                center=np.array([o["center"] for o in objs], np.float32).reshape(-1, 2),
                fullbox=np.array([o["fullbox"] for o in objs], np.float32).reshape(-1, 4))


def gen_to3(r, seed):
    """Real to_3channel (common/utils/image.py:72-100) for all flag combinations."""
    from numba.typed import List
    cols = [(32, 32, 64), (0, 0, 255), (96, 128, 128), (102, 255, 0), (255, 0, 204)]   # label_spec.py:23-30
    items = List([(k, v) for k, v in zip("abcde", cols)])
    rng = np.random.default_rng(seed)
    raw = rng.uniform(0, 1, (24, 40, 7)).astype(np.float32)
    raw[0, 0, :5] = 0.3                          # all equal -> NaN under apply_softmax
    raw[0, 1, :5] = [.1, .5, .5, .2, 0]          # tie -> first index
    raw[0, 2, :5] = [-.5, -.2, -.9, -.3, -.4]    # negatives
    raw[0, 3, :5] = [.2, 1.7, .1, .3, .4]        # score clipped to 1
    out = dict(raw=raw, colours=np.array(cols, np.int32))
    for thr in (None, 0.3):
        for uw in (False, True):
            for sm in (True, False):
                key = f"out_thr{'N' if thr is None else '1'}_w{int(uw)}_s{int(sm)}"
                out[key] = r["to_3channel"](raw.copy(), items, thr, uw, sm)
    return out


def main():
    r = ref_import.load()
    np.savez_compressed(os.path.join(HERE, "render_process_a.npz"), **gen_render_process(r, 11, 14, 192, 96))
    np.savez_compressed(os.path.join(HERE, "render_process_b.npz"), **gen_render_process(r, 12, 40, 256, 128))
    np.savez_compressed(os.path.join(HERE, "render_process_empty.npz"), **gen_render_process(r, 13, 0, 64, 32))
    np.savez_compressed(os.path.join(HERE, "fill_cases.npz"), **gen_fill_cases(r, 21))
    np.savez_compressed(os.path.join(HERE, "decode_r.npz"), **gen_decode_r(r, 31))
    np.savez_compressed(os.path.join(HERE, "to3.npz"), **gen_to3(r, 41))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
