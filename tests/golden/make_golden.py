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


def gen_render_process_3d(r, seed, n_obj, in_w, in_h):
    """Real ProcessImages.process with the l_shape and 3d_info fields active (processor.py:69-115,283-299): objects with a
    projected 3D box (8 corners), some without (-> ignore areas), some with a far-off / non-convex projection."""
    import cv2
    P = r["CenternetParams"](6)
    P.INPUT_WIDTH, P.INPUT_HEIGHT = in_w, in_h
    P.REGRESSION_FIELDS["l_shape"].active = True
    P.REGRESSION_FIELDS["3d_info"].active = True
    proc = r["ProcessImages"](P)
    ok, buf = cv2.imencode(".png", np.zeros((in_h, in_w, 3), np.uint8))
    rng = np.random.default_rng(seed)
    boxes, cls, valid, box3d, info = [], [], [], [], []
    for i in range(n_obj):
        w = float(np.exp(rng.uniform(np.log(6), np.log(in_w / 3))))
        h = float(np.exp(rng.uniform(np.log(6), np.log(in_h / 2))))
        x, y = float(rng.uniform(-10, in_w - 8)), float(rng.uniform(-8, in_h - 8))
        boxes.append([x, y, w, h])
        cls.append(int(rng.integers(0, 6)))
        valid.append(i % 5 != 2)
        # 8 corners: front face (0..3) and back face (4..7), top = 0,3,4,7 / bottom = 1,2,5,6 (processor.py:76-77)
        dx, dy = rng.uniform(-0.3, 0.3) * w, rng.uniform(-0.15, 0.15) * h
        front = [[x, y], [x, y + h], [x + 0.7 * w, y + h * rng.uniform(0.85, 1.1)], [x + 0.7 * w, y]]
        back = [[x + dx + 0.3 * w, y + dy], [x + dx + 0.3 * w, y + dy + 0.8 * h], [x + dx + w, y + dy + 0.8 * h], [x + dx + w, y + dy]]
        pts = np.array(front + back, np.float64)
        if i % 7 == 4:
            pts += 4 * in_w                                  # projection far outside: invalid L-shape (:107-115)
        if i % 7 == 5:
            pts[[2, 5, 6], 1] = y + h * 1.2                   # the middle bottom point above both others: the non-convex branch (:90-93)
            pts[1, 1] = y + h * 1.3
        box3d.append(pts.reshape(-1))
        info.append([float(rng.uniform(-20, 20)), float(rng.uniform(-2, 2)), float(rng.uniform(3, 80)), float(rng.uniform(-3.1, 3.1)),
                     float(rng.uniform(1.4, 2.6)), float(rng.uniform(1.2, 3.5)), float(rng.uniform(2.5, 12))])
    objs = [{"box2d": b, "obj_class": NAMES[c], "box3d_valid": bool(v), "box3d": list(map(float, k)), "x": t[0], "y": t[1], "z": t[2],
             "orientation": t[3], "width": t[4], "height": t[5], "length": t[6]}
            for b, c, v, k, t in zip(boxes, cls, valid, box3d, info)]
    _, _, gt, _ = proc.process({"img": buf.tobytes(), "objects": objs}, None, None, {"epoch": 0})
    return dict(raw_boxes=np.array(boxes, np.float64), cls=np.array(cls, np.int32), valid=np.array(valid), box3d=np.array(box3d, np.float64),
                info=np.array(info, np.float64), y_true=gt.astype(np.float32), in_w=in_w, in_h=in_h)


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


# ---- loss: the UNMODIFIED reference loss files executed over oracle/tf_shim.py (TF ops on torch-CPU fp32) ----------------
from oracle.ref_harness import ref_loss_object  # noqa: E402,F401


def ref_loss_values(loss, y_true, y_pred):
    """Every public term of the reference loss on one (y_true with weights plane, y_pred) pair -> dict of python floats."""
    import torch
    yt = torch.from_numpy(np.ascontiguousarray(y_true, np.float32))
    yp = torch.from_numpy(np.ascontiguousarray(y_pred, np.float32))
    out = {"total": float(loss(y_true, y_pred)),                                            # Loss.__call__ -> call, loss.py:133-155
           "focal_weighted": float(loss.obj_focal_loss(yt[..., :-1], yp, yt[..., -1])),     # as call uses it, :137-140
           "focal_metric": float(loss.obj_focal_loss(yt, yp))}                              # as a Keras metric, train.py:62
    for name in ("class", "r_offset", "fullbox", "l_shape", "radial_dist", "orientation", "obj_dims", "track_offset"):
        pos = {"radial_dist": "radial_dist_pos", "orientation": "orientation_pos", "obj_dims": "obj_dims_pos"}.get(name, name + "_pos")
        if hasattr(loss, pos):
            out[name] = float(getattr(loss, name + "_loss")(yt[..., :-1], yp))
    return out


def _pack_cases(cases):
    """cases: {name: (y_true, y_pred, values dict)} -> flat dict for np.savez (pred stored as given dtype)."""
    out = {"cases": np.array(sorted(cases))}
    for name, (yt, yp, vals) in cases.items():
        out[name + "/y_true"] = yt
        out[name + "/y_pred"] = yp
        out[name + "/terms"] = np.array(sorted(vals))
        out[name + "/values"] = np.array([vals[k] for k in sorted(vals)], np.float64)
    return out


def gen_loss_fixture(r):
    """The reference's own test inputs (loss_test.py:9-49, centertracker/loss_test.py:10-21 with the track channels in
    the intended place, before the weights plane) and the perturbations its tests apply."""
    sys.path.insert(0, os.path.dirname(HERE))
    from fixtures import reference_loss_fixture
    cases = {}
    for track in (False, True):
        nb, gt, perfect = reference_loss_fixture(track=track)
        loss, P = ref_loss_object(r, nb, track=track, l_shape=True, info3d=True)
        assert P.mask_channels() + 1 == gt.shape[-1]
        tag = "trk_" if track else ""
        preds = {"perfect": perfect}
        p = perfect.copy(); p[0, 1, 1, 0] = 0.8; preds["peak08"] = p               # loss_test.py:57
        p = perfect.copy(); p[0, 2, 1, 0] = 1.0; preds["one_off"] = p              # :62
        p = perfect.copy(); p[0, 6, 1, 0] = 1.0; preds["wrong_peak"] = p           # :65-66
        p = perfect.copy(); p[0, 1, 1, 2] = 0.8; preds["class08"] = p              # :74
        p = perfect.copy(); p[0, 1, 1, 3] = 1.0; preds["class_wrong"] = p          # :79
        p = perfect.copy(); p[0, 1, 1, 6] = 2.2 - 30; preds["box_far"] = p         # :89
        p = perfect.copy(); p[0, 1, 1, 5:20] += np.linspace(-1.5, 2.5, 15).astype(np.float32); preds["all_fields_off"] = p
        if track:
            p = perfect.copy(); p[0, 1, 1, 20:22] = [0.0, 3.0]; preds["track_off"] = p   # centertracker/loss_test.py:28-30
        for k, pr in preds.items():
            cases[tag + k] = (gt, pr, ref_loss_values(loss, gt, pr))
    return _pack_cases(cases)


def _synth_batch(nb, profile, H, W, B, config, track=False, l_shape=False, info3d=False, n_obj=None):
    from oracle import render_np
    from oracle.layout import make_layout
    import synth
    Lo = make_layout(H, W, nb, profile, track=track, l_shape=l_shape, info3d=info3d)
    data = synth.make_batch(Lo, config, B, n_obj=n_obj, track=track)
    yt = np.stack([render_np.render_image(Lo, data["boxes"][i], data["cls"][i], data["ignore"][i],
                                          data["track"][i] if track else None) for i in range(B)])
    return Lo, yt, data["y_pred"]


def gen_loss_maps(r):
    """Full-size and odd-size maps: y_true from the render oracle (itself pinned bit-exact to the real render), y_pred
    synthetic (SURVEY 8d distributions) rounded to fp16-representable values so the fixture stays small."""
    sys.path.insert(0, os.path.dirname(HERE))
    cases = {}
    rng = np.random.default_rng(77)
    for name, nb, profile, H, W, B, kw in (
            ("profile_r_128x384", 10, "R", 128, 384, 1, {}),
            ("profile_n_128x384", 10, "N", 128, 384, 1, {}),
            ("tracker_n_64x96", 10, "N", 64, 96, 2, dict(track=True)),
            ("tracker_r_37x53", 6, "R", 37, 53, 3, dict(track=True)),
            ("all_fields_r_32x48", 4, "R", 32, 48, 3, dict(track=True, l_shape=True, info3d=True))):
        Lo, yt, yp = _synth_batch(nb, profile, H, W, B, 7, **kw)
        # targets for the fields the render does not fill (l_shape / 3d_info): random values at the peak pixels
        pm = (yt[..., :Lo.hm] == 1.0).any(-1)
        for f in Lo.fields:
            if f.name in ("l_shape", "radial_dist", "orientation", "obj_dims"):
                yt[..., f.off:f.off + f.size][pm] = rng.normal(0, 3, (int(pm.sum()), f.size)).astype(np.float32)
        yp = yp.astype(np.float16)
        loss, P = ref_loss_object(r, nb, hm=Lo.hm, **kw)
        assert P.mask_channels() + (Lo.hm - 1) == Lo.Cp, (P.mask_channels(), Lo.Cp)
        cases[name] = (yt, yp, ref_loss_values(loss, yt, yp.astype(np.float32)))
    # no object anywhere: the n == 0 branches of tf.cond (loss.py:59,130)
    Lo, yt, yp = _synth_batch(5, "R", 16, 24, 2, 8, n_obj=0)
    yt[..., 0] = rng.uniform(0, 0.9, yt.shape[:-1]).astype(np.float32)
    loss, _ = ref_loss_object(r, 5)
    yp = yp.astype(np.float16)
    cases["no_objects_r_16x24"] = (yt, yp, ref_loss_values(loss, yt, yp.astype(np.float32)))
    # clip edges of the log arguments and exact 0 / 1 predictions (loss.py:41,47)
    Lo, yt, yp = _synth_batch(3, "N", 9, 11, 2, 9)
    yp[0, 0, 0, 0], yp[0, 0, 1, 0], yp[0, 0, 2, 0], yp[0, 0, 3, 0] = 0.0, 1.0, 0.01, 0.99
    yp[1, 4, 4, :3] = [0.005, 0.995, 0.5]
    yp = yp.astype(np.float16)
    loss, _ = ref_loss_object(r, 3, hm=3)
    cases["clip_edges_n_9x11"] = (yt, yp, ref_loss_values(loss, yt, yp.astype(np.float32)))
    return _pack_cases(cases)


def gen_loss_multitask(r):
    """MultitaskLoss.calc_centernet (models/multitask/loss.py:44-47): the CenterNet slice of the wide multitask tensors."""
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    nb = 7
    mp = r["MultitaskParams"](nb)
    ml = r["MultitaskLoss"](mp)
    Ccn = mp.cn_params.mask_channels()
    H, W, B = 24, 40, 2
    Lo, yt, yp = _synth_batch(nb, "R", H, W, B, 10)
    assert Lo.Cp == Ccn
    rng = np.random.default_rng(78)
    ct, cp = ml.depth_offset["y_true"][1], ml.depth_offset["y_pred"][1]
    wide_t = rng.normal(0, 1, (B, H, W, ct)).astype(np.float32)
    wide_p = rng.normal(0, 1, (B, H, W, cp)).astype(np.float16)
    wide_t[..., :Ccn + 1] = yt
    wide_p[..., :Ccn] = yp.astype(np.float16)
    v = float(ml.calc_centernet(torch.from_numpy(wide_t), torch.from_numpy(wide_p.astype(np.float32))))
    return dict(y_true=wide_t, y_pred=wide_p, nb_classes=nb, cn_channels=Ccn, total=np.float64(v),
                cn_offset_true=np.array(ml.cn_offset["y_true"]), cn_offset_pred=np.array(ml.cn_offset["y_pred"]),
                semseg_offset_pred=np.array(ml.semseg_offset["y_pred"]), depth_offset_pred=np.array(ml.depth_offset["y_pred"]))


def main():
    r = ref_import.load()
    np.savez_compressed(os.path.join(HERE, "loss_fixture.npz"), **gen_loss_fixture(r))
    np.savez_compressed(os.path.join(HERE, "loss_maps.npz"), **gen_loss_maps(r))
    np.savez_compressed(os.path.join(HERE, "loss_multitask.npz"), **gen_loss_multitask(r))
    np.savez_compressed(os.path.join(HERE, "render_process_a.npz"), **gen_render_process(r, 11, 14, 192, 96))
    np.savez_compressed(os.path.join(HERE, "render_process_b.npz"), **gen_render_process(r, 12, 40, 256, 128))
    np.savez_compressed(os.path.join(HERE, "render_process_empty.npz"), **gen_render_process(r, 13, 0, 64, 32))
    np.savez_compressed(os.path.join(HERE, "render_process_3d.npz"), **gen_render_process_3d(r, 14, 30, 256, 128))
    np.savez_compressed(os.path.join(HERE, "fill_cases.npz"), **gen_fill_cases(r, 21))
    np.savez_compressed(os.path.join(HERE, "decode_r.npz"), **gen_decode_r(r, 31))
    np.savez_compressed(os.path.join(HERE, "to3.npz"), **gen_to3(r, 41))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
