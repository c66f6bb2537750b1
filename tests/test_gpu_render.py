"""GPU parity: cvm_render_gt / cvm_render_prev_hm / cvm_fill_heatmap_inplace (through the C ABI) against the golden
vectors produced by the REAL reference functions and against the CPU oracle on seeded inputs.
Tolerance: heatmaps within 1e-5 relative (north_star); in practice the fp64 device math is bit-identical."""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import render_np
from oracle.layout import make_layout

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _params(nb, per_class, H, W, track=False):
    from cvmhot.models.centernet import CenternetParams
    from cvmhot.models.centertracker import CentertrackerParams
    p = (CentertrackerParams if track else CenternetParams)(nb, per_class)
    p.INPUT_HEIGHT, p.INPUT_WIDTH = H * p.R, W * p.R
    return p


def _close(a, b):
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=1e-30)


@pytest.mark.parametrize("name", ["render_process_a", "render_process_b", "render_process_empty"])
def test_render_vs_real_process_golden(cuda, golden_dir, name):
    from cvmhot.models.centernet import ProcessImages
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    names = ["car", "truck", "van", "motorbike", "cyclist", "ped"]
    p = _params(6, False, int(g["in_h"]) // 2, int(g["in_w"]) // 2)
    proc = ProcessImages(p)
    objs = [{"box2d": list(b), "obj_class": names[c]} for b, c in zip(g["raw_boxes"], g["cls"])]
    y = proc.render_batch([{"objects": objs}])[0].cpu().numpy()
    assert y.shape == g["y_true"].shape
    _close(y, g["y_true"])
    assert (y == g["y_true"]).mean() == 1.0     # bit-exact against the reference's own output
    # the per-sample plug-in entry point gives the same thing as numpy
    img = np.zeros((int(g["in_h"]), int(g["in_w"]), 3), np.uint8)
    _, inp, gt, _ = proc.process({"img": img, "objects": objs}, None, None, {"epoch": 0})
    assert inp[0].dtype == np.float32 and np.array_equal(gt, y)


def test_fill_heatmap_dropin_vs_golden(cuda, golden_dir):
    from cvmhot.models.centernet import fill_heatmap
    g = np.load(os.path.join(golden_dir, "fill_cases.npz"))
    H, W = int(g["H"]), int(g["W"])
    heat = np.zeros((H, W, 1), np.float32)
    wts = np.ones((H, W), np.float32)
    for cx, cy, w, h, peak in g["recs"]:
        fill_heatmap(heat, 0.9, 2, wts, int(cx), int(cy), w, h, W, H, peak)
    _close(heat, g["heat"])
    _close(wts, g["wts"])


def test_prev_heatmap_vs_golden(cuda, golden_dir):
    from cvmhot.models.centertracker import CenterTrackerProcess
    g = np.load(os.path.join(golden_dir, "fill_cases.npz"))
    H, W = int(g["H"]), int(g["W"])
    p = _params(6, False, H, W, track=True)
    proc = CenterTrackerProcess(p)
    recs = [(int(r[0]), int(r[1]), r[2], r[3], r[4]) for r in g["recs"]]
    out = proc.render_prev_batch([recs, [], recs[:5]])
    assert out.shape == (3, H, W, 1)
    _close(out[0].cpu().numpy(), g["heat"])
    assert float(out[1].abs().max()) == 0.0
    _close(out[2].cpu().numpy(), render_np.render_prev_heatmap(H, W, 0.9, 2, recs[:5]))


@pytest.mark.parametrize("profile,H,W,K,track", [("N", 128, 384, 10, False), ("R", 128, 384, 10, False),
                                                 ("N", 128, 384, 10, True), ("R", 9, 11, 3, False),
                                                 ("N", 37, 53, 7, False), ("N", 64, 1000, 4, False),
                                                 ("R", 9, 11, 130, False)])   # (136-float pixels: the scalar fill)
def test_render_vs_oracle(cuda, profile, H, W, K, track):
    from cvmhot.models.centernet.processor import ProcessImages, pack_objects, pack_boxes
    from cvmhot.layout import layout_from_params
    p = _params(K, profile == "N", H, W, track)
    Lo = make_layout(H, W, K, profile, track=track)
    L = layout_from_params(p)
    assert (L.Cp, L.Ct, L.off_roff, L.off_box, L.off_track, L.off_class) == (Lo.Cp, Lo.Ct, Lo.off_roff, Lo.off_box, Lo.off_track, Lo.off_class)
    B = 4
    data = synth.make_batch(Lo, 2, B, track=track)
    data["boxes"][1] = np.zeros((0, 4))            # an image without objects
    data["cls"][1] = np.zeros((0,), np.int32)
    if track:
        data["track"][1] = np.zeros((0, 2), np.float32)
    if len(data["boxes"][0]) >= 2:                   # two objects on one centre pixel, different classes
        data["boxes"][0][1] = data["boxes"][0][0] + [0.25, 0.25, 0, 0]
        data["cls"][0][1] = (data["cls"][0][0] + 1) % K
    proc = ProcessImages(p) if not track else __import__("cvmhot.models.centertracker", fromlist=["x"]).CenterTrackerProcess(p)
    rec, offs = pack_objects(data["boxes"], data["cls"], data["track"])
    y = proc.render_packed(L, rec, offs, *pack_boxes(data["ignore"])).cpu().numpy()
    for b in range(B):
        ref = render_np.render_image(Lo, data["boxes"][b], data["cls"][b], data["ignore"][b],
                                     data["track"][b] if track else None)
        _close(y[b], ref)
        assert (y[b][..., :Lo.hm] == 1.0).sum() == (ref[..., :Lo.hm] == 1.0).sum()   # the loss's ==1.0 test depends on exact peaks


def test_render_many_objects_and_chunks(cuda):
    """> 64 objects per image exercises the chunked object loop and the cross-chunk 'last writer wins' rule."""
    from cvmhot.models.centernet.processor import ProcessImages, pack_objects, pack_boxes
    from cvmhot.layout import layout_from_params
    H, W, K = 64, 96, 5
    p = _params(K, True, H, W)
    Lo = make_layout(H, W, K, "N")
    rng = np.random.default_rng(5)
    boxes, cls = synth.objects_for_image(rng, H, W, 2, 150, K)
    boxes[140] = boxes[3]          # same centre pixel, 137 objects apart
    boxes[140][2:] = boxes[3][2:] * 0.5
    boxes[140][:2] = boxes[3][:2] + boxes[3][2:] * 0.25
    rec, offs = pack_objects([boxes], [cls])
    y = ProcessImages(p).render_packed(layout_from_params(p), rec, offs, *pack_boxes([np.zeros((0, 4))])).cpu().numpy()
    _close(y[0], render_np.render_image(Lo, boxes, cls, []))


def test_render_properties_full_size(cuda):
    """BASELINE configs[1] shape (B sampled down to 48): size-independent properties of the render, plus parity of a few
    images against the oracle.  Every object centre holds exactly 1.0 in its class plane (the loss's `== 1` test), heat
    lies in [0, 1], weights are <= 1 and exactly 0 inside ignore boxes, regression channels are zero away from centres,
    and rendering is idempotent / independent of the batch an image sits in."""
    import bench
    from cvmhot import ops
    from cvmhot.layout import layout_from_params
    from cvmhot.models.centernet.processor import pack_boxes, pack_objects
    H, W, K, B = 128, 384, 10, 48
    p = _params(K, True, H, W)
    L = layout_from_params(p)
    Lo = make_layout(H, W, K, "N")
    boxes, cls, ign = bench.gen_objects(0, B)
    rec, offs = pack_objects(list(boxes), list(cls))
    irec, ioffs = pack_boxes(list(ign))
    dev = cuda
    args = (ops.to_device_records(rec, ops.OBJ_DTYPE, dev), torch.from_numpy(offs).to(dev), B,
            ops.to_device_records(irec, ops.BOX_DTYPE, dev), torch.from_numpy(ioffs).to(dev))
    y = ops.render_gt(L, *args)
    y2 = ops.render_gt(L, *args)
    assert torch.equal(y, y2)                                                     # deterministic
    heat, wts = y[..., :K], y[..., -1]
    assert float(heat.min()) >= 0.0 and float(heat.max()) == 1.0 and float(wts.max()) <= 1.0
    cx, cy = bench.centres(boxes)
    bi = np.repeat(np.arange(B), boxes.shape[1])
    at_centres = heat[torch.from_numpy(bi).to(dev), torch.from_numpy(cy.reshape(-1)).to(dev),
                      torch.from_numpy(cx.reshape(-1)).to(dev), torch.from_numpy(cls.reshape(-1).astype(np.int64)).to(dev)]
    big_enough = torch.from_numpy((boxes[..., 2] >= 2).reshape(-1) & (boxes[..., 3] >= 2).reshape(-1)).to(dev)
    assert bool((at_centres[big_enough] == 1.0).all())                            # exact peaks (objects < 2 px draw nothing)
    reg = y[..., K:-1].abs().sum(dim=-1)
    mask = torch.zeros((B, H, W), dtype=torch.bool, device=dev)
    mask[torch.from_numpy(bi).to(dev), torch.from_numpy(cy.reshape(-1)).to(dev), torch.from_numpy(cx.reshape(-1)).to(dev)] = True
    assert float(reg[~mask].max()) == 0.0                                         # targets only at centre pixels
    for b in (0, 17, 47):
        for (x, yy, w, h) in ign[b]:
            sl = wts[b, max(int(yy), 0):int(yy + h), max(int(x), 0):int(x + w)]
            assert sl.numel() == 0 or float(sl.abs().max()) == 0.0                # ignore areas (input px used as mask px)
        ref = render_np.render_image(Lo, boxes[b], cls[b], ign[b])
        _close(y[b].cpu().numpy(), ref)
    # an image renders the same alone as inside the batch
    r1, o1 = pack_objects([boxes[17]], [cls[17]])
    i1, io1 = pack_boxes([ign[17]])
    y17 = ops.render_gt(L, ops.to_device_records(r1, ops.OBJ_DTYPE, dev), torch.from_numpy(o1).to(dev), 1,
                        ops.to_device_records(i1, ops.BOX_DTYPE, dev), torch.from_numpy(io1).to(dev))
    assert torch.equal(y17[0], y[17])


def test_device_front_end_vs_real_process_golden(cuda, golden_dir):
    """cvm_prepare_objects + cvm_render_gt on RAW (unclipped, unfiltered) boxes reproduce, bit for bit, the y_true the real
    ProcessImages.process produced from the same raw boxes; in a batch with an empty image in the middle too; and the
    records equal the host mirror of the filter."""
    from cvmhot import ops
    from cvmhot.models.centernet import ProcessImages
    from cvmhot.models.centernet.processor import pack_boxes, pack_objects
    for name in ("render_process_a", "render_process_b", "render_process_empty"):
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        p = _params(6, False, int(g["in_h"]) // 2, int(g["in_w"]) // 2)
        proc = ProcessImages(p)
        raw1 = g["raw_boxes"].reshape(-1, 4).astype(np.float64)
        cls1 = g["cls"].reshape(-1).astype(np.int32)
        n1 = len(cls1)
        # batch: the sample, an image without objects, the sample again
        raw = np.concatenate([raw1, raw1]) if n1 else np.zeros((0, 4))
        cls = np.concatenate([cls1, cls1]) if n1 else np.zeros((0,), np.int32)
        offs = np.array([0, n1, n1, 2 * n1], np.int32)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
        y = proc.render_raw_batch(dev(raw.reshape(-1, 4)), dev(cls), dev(offs))
        y = y.cpu().numpy()
        assert np.array_equal(y[0], g["y_true"]) and np.array_equal(y[2], g["y_true"])
        assert float(np.abs(y[1][..., :-1]).max()) == 0.0 and float(y[1][..., -1].min()) == 1.0
        if n1 == 0:
            continue
        # record-level check against the host mirror of the filter
        objs_d, offs_d, ign_d, ioffs_d = ops.prepare_objects(dev(raw), dev(cls), dev(offs), p.INPUT_WIDTH, p.INPUT_HEIGHT,
                                                             p.MIN_BOX_AREA)
        boxes, c, ign = proc.filter_objects([{"box2d": list(bx), "obj_class": int(k)} for bx, k in zip(raw1, cls1)],
                                            p.INPUT_WIDTH, p.INPUT_HEIGHT)
        rec, roffs = pack_objects([boxes, [], boxes], [c, [], c])
        irec, ioffs = pack_boxes([ign, [], ign])
        assert np.array_equal(offs_d.cpu().numpy(), roffs) and np.array_equal(ioffs_d.cpu().numpy(), ioffs)
        got = objs_d.cpu().numpy()[:rec.size * ops.OBJ_DTYPE.itemsize].view(ops.OBJ_DTYPE)
        for f in ("x", "y", "w", "h", "cls"):
            assert np.array_equal(got[f], rec[f]), f
        goti = ign_d.cpu().numpy()[:irec.size * ops.BOX_DTYPE.itemsize].view(ops.BOX_DTYPE)
        for f in ("x", "y", "w", "h"):
            assert np.array_equal(goti[f], irec[f]), f


def test_render_l_shape_and_3d_info_vs_real_process_golden(cuda, golden_dir):
    """l_shape (7) and 3d_info (5) targets active (reference processor.py:69-115,283-299): objects without a valid L-shape
    turn into ignore areas, the non-convex projection takes the box height; y_true equals the REAL process() output."""
    from cvmhot.models.centernet import ProcessImages
    g = np.load(os.path.join(golden_dir, "render_process_3d.npz"))
    names = ["car", "truck", "van", "motorbike", "cyclist", "ped"]
    p = _params(6, False, int(g["in_h"]) // 2, int(g["in_w"]) // 2)
    p.REGRESSION_FIELDS["l_shape"].active = True
    p.REGRESSION_FIELDS["3d_info"].active = True
    assert p.mask_channels() + 1 == g["y_true"].shape[-1]
    objs = [{"box2d": list(b), "obj_class": names[c], "box3d_valid": bool(v), "box3d": list(k), "x": t[0], "y": t[1], "z": t[2],
             "orientation": t[3], "width": t[4], "height": t[5], "length": t[6]}
            for b, c, v, k, t in zip(g["raw_boxes"], g["cls"], g["valid"], g["box3d"], g["info"])]
    proc = ProcessImages(p)
    y = proc.render_batch([{"objects": objs}, {"objects": objs[:7]}])
    y0 = y[0].cpu().numpy()
    assert (y0 == g["y_true"]).mean() == 1.0
    _, _, gt, _ = proc.process({"img": np.zeros((int(g["in_h"]), int(g["in_w"]), 3), np.uint8), "objects": objs}, None, None, {"epoch": 0})
    assert np.array_equal(gt, g["y_true"])
    # the loss mirror consumes these targets (all fields active)
    from cvmhot.models.centernet import CenternetLoss
    yp = torch.rand((2,) + tuple(y.shape[1:3]) + (p.mask_channels(),), device=cuda)
    assert torch.isfinite(CenternetLoss(p)(y, yp))


@pytest.mark.parametrize("H,W,B", [(16, 24, 1500), (40, 56, 700)])
def test_render_many_small_images_per_cta(cuda, H, W, B):
    """Thousands of small images: every persistent CTA walks through many images, so the setup group's double-buffered
    table sets rotate constantly while the builder groups sit in different images.  The batch must equal the same images
    rendered a few at a time (bitwise), and a sample must match the oracle."""
    from cvmhot.models.centernet.processor import ProcessImages, pack_objects, pack_boxes
    from cvmhot.layout import layout_from_params
    K = 4
    p = _params(K, True, H, W)
    Lo = make_layout(H, W, K, "N")
    L = layout_from_params(p)
    rng = np.random.default_rng(77)
    boxes, cls, ign = [], [], []
    for b in range(B):
        n = int(rng.integers(0, 9)) if b % 7 else 0          # images without objects in between
        if b % 50 == 3:
            n = 80                                             # ... and crowded ones (more objects than the tables hold)
        bx, c = synth.objects_for_image(rng, H, W, 2, n, K)
        boxes.append(bx)
        cls.append(c)
        ign.append(synth.ignore_for_image(rng, H, W, 20) if b % 60 == 0 else
                   np.array([[rng.integers(0, W), rng.integers(0, H), 5, 4]], np.float64) if b % 5 == 0 else np.zeros((0, 4)))
    proc = ProcessImages(p)
    rec, offs = pack_objects(boxes, cls)
    y = proc.render_packed(L, rec, offs, *pack_boxes(ign)).cpu().numpy()
    step = 97
    for s in range(0, B, step):
        r2, o2 = pack_objects(boxes[s:s + step], cls[s:s + step])
        part = proc.render_packed(L, r2, o2, *pack_boxes(ign[s:s + step])).cpu().numpy()
        assert np.array_equal(y[s:s + step], part), s
    for b in list(range(0, B, 131)) + [3, 53, 60, B - 1]:
        _close(y[b], render_np.render_image(Lo, boxes[b], cls[b], ign[b]))


def test_render_repeatable_bitwise(cuda):
    """The builder groups and the setup group of a CTA only meet through flags and named barriers: a missing hand-over would
    show up as run-to-run differences.  40 launches of the BASELINE shape (96 images: CTAs with one, two and three images)
    must be bitwise identical, also while another stream keeps the GPU busy."""
    import bench
    from cvmhot import ops
    from cvmhot.layout import layout_from_params
    from cvmhot.models.centernet.processor import pack_boxes, pack_objects
    B = 96
    p = _params(10, True, 128, 384)
    L = layout_from_params(p)
    boxes, cls, ign = bench.gen_objects(0, B)
    rec, offs = pack_objects(list(boxes), list(cls))
    ign_rec, ign_offs = pack_boxes(list(ign))
    objs_d = ops.to_device_records(rec, ops.OBJ_DTYPE, cuda)
    offs_d = torch.from_numpy(offs).to(cuda)
    ign_d = ops.to_device_records(ign_rec, ops.BOX_DTYPE, cuda)
    ioffs_d = torch.from_numpy(ign_offs).to(cuda)
    ref = ops.render_gt(L, objs_d, offs_d, B, ign_d, ioffs_d).clone()
    noise = torch.empty(2 ** 26, device=cuda)
    side = torch.cuda.Stream(device=cuda)
    out = torch.empty_like(ref)
    for it in range(40):
        if it % 2:
            with torch.cuda.stream(side):
                noise.fill_(float(it))
        ops.render_gt(L, objs_d, offs_d, B, ign_d, ioffs_d, out=out)
        assert torch.equal(out, ref), it
    torch.cuda.synchronize()
