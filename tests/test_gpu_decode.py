"""GPU parity: cvm_decode_topk (canonical CenterNet decode) and cvm_decode_window9 (the reference's process_2d_output)
through the C ABI.  Bar (north_star): top-K flat indices, class ids and scores BIT-EXACT under the lowest-flat-index
tie-break; boxes/centres are fp32 op-for-op like NumPy, so they are compared exactly too."""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import decode_np
from oracle.layout import make_layout

pytestmark = pytest.mark.gpu


def _params(nb, per_class, H, W, track=False):
    from cvmhot.models.centernet import CenternetParams
    from cvmhot.models.centertracker import CentertrackerParams
    p = (CentertrackerParams if track else CenternetParams)(nb, per_class)
    p.INPUT_HEIGHT, p.INPUT_WIDTH = H * p.R, W * p.R
    return p


def _check_topk(out, ref, K_eff):
    for k in ("scores", "cls", "flat", "centers", "boxes"):
        got = out[k].cpu().numpy()[:, :K_eff]
        assert np.array_equal(got, ref[k]), k


@pytest.mark.parametrize("H,W,K_cls,track,B,K", [(128, 384, 10, False, 4, 100), (128, 384, 10, True, 3, 100),
                                                 (37, 53, 7, False, 5, 100), (64, 1000, 4, False, 2, 50),
                                                 (9, 11, 1, False, 2, 100), (20, 20, 3, False, 3, 7)])
def test_topk_vs_oracle(cuda, H, W, K_cls, track, B, K):
    from cvmhot.models.centernet.post_processing import decode_topk
    from cvmhot.common.utils import Roi
    Lo = make_layout(H, W, K_cls, "N", track=track)
    data = synth.make_batch(Lo, 4, B, track=track)
    yp = data["y_pred"]
    rois = [(1.0, 0, 0), (0.25, -5, -3), (0.5, 3, 7), (2.0, 0, 1), (1.0, 0, 0)][:B]
    ref = decode_np.decode_topk(Lo, yp, K, rois)
    roi_objs = []
    for s, l, t in rois:
        r = Roi(); r.scale, r.offset_left, r.offset_top = s, l, t
        roi_objs.append(r)
    out = decode_topk(torch.from_numpy(yp).to(cuda), _params(K_cls, True, H, W, track), K=K, rois=roi_objs)
    K_eff = ref["scores"].shape[1]
    _check_topk(out, ref, K_eff)
    if track:
        assert np.array_equal(out["track"].cpu().numpy()[:, :K_eff], ref["track"])
    if K_eff < K:      # fewer elements than K: the rest is marked invalid
        assert (out["flat"].cpu().numpy()[:, K_eff:] == -1).all() and (out["cls"].cpu().numpy()[:, K_eff:] == -1).all()


def test_topk_ties_plateaus_and_sparse_maps(cuda):
    from cvmhot.models.centernet.post_processing import decode_topk
    H, W, C = 48, 64, 6
    Lo = make_layout(H, W, C, "N")
    rng = np.random.default_rng(11)
    yp = np.zeros((5, H, W, Lo.Cp), np.float32)
    yp[..., Lo.hm:] = rng.uniform(0, 50, (5, H, W, Lo.Cp - Lo.hm))
    # 0: only 7 positive peaks -> tail of score-0 entries in flat order (some of the first flats are peaks themselves)
    for (y, x, c, v) in [(0, 0, 0, .5), (0, 0, 2, .7), (0, 1, 1, .7), (10, 10, 3, .2), (47, 63, 5, .9), (20, 5, 0, .9), (0, 3, 4, .1)]:
        yp[0, y, x, c] = v
    # 1: constant map -> everything is a plateau peak with equal score: the first K flat indices win
    yp[1, ..., :C] = 0.3
    # 2: all zeros -> no positive peak at all
    # 3: many exact duplicates of a few values (heavy ties across channels and rows)
    yp[3, ..., :C] = rng.choice(np.array([0.1, 0.2, 0.7, 0.9], np.float32), size=(H, W, C))
    # 4: monotone ramp (every row has exactly one peak per channel at the right border) + a negative region
    yp[4, ..., :C] = (np.arange(W, dtype=np.float32)[None, :, None] / W + 0.001 * np.arange(C, dtype=np.float32)) - 0.2
    ref = decode_np.decode_topk(Lo, yp, 100)
    out = decode_topk(torch.from_numpy(yp).to(cuda), _params(C, True, H, W), K=100)
    _check_topk(out, ref, 100)
    assert (ref["scores"][2] == 0).all() and np.array_equal(ref["flat"][2], np.arange(100))
    assert np.array_equal(ref["flat"][1], np.arange(100))


def test_topk_compaction_stress(cuda):
    """Dense, slowly rising maps force many threshold raises and compactions inside one segment."""
    from cvmhot.models.centernet.post_processing import decode_topk
    H, W, C = 128, 384, 10
    Lo = make_layout(H, W, C, "N")
    rng = np.random.default_rng(12)
    yp = np.zeros((2, H, W, Lo.Cp), np.float32)
    base = rng.uniform(0.01, 0.5, (2, H, W, C)).astype(np.float32)
    ramp = (np.arange(H, dtype=np.float32) / H * 0.45)[None, :, None, None]      # later rows score higher
    yp[..., :C] = base + ramp
    yp[1, ..., :C] = yp[1, ::-1, :, :C]                                             # and the reverse
    ref = decode_np.decode_topk(Lo, yp, 100)
    out = decode_topk(torch.from_numpy(yp).to(cuda), _params(C, True, H, W), K=100)
    _check_topk(out, ref, 100)


def test_topk_properties_full_size(cuda):
    """BASELINE config 2 shape (B=256 is sampled down to 64 to keep the host reference affordable):
    size-independent properties + spot parity on a few images."""
    from cvmhot.models.centernet.post_processing import decode_topk
    H, W, C, B = 128, 384, 10, 64
    Lo = make_layout(H, W, C, "N")
    g = torch.Generator(device="cuda").manual_seed(5)
    yp = torch.empty((B, H, W, Lo.Cp), device=cuda)
    yp[..., :C] = torch.sigmoid(torch.randn((B, H, W, C), device=cuda, generator=g) * 1.5 - 4.0)
    yp[..., C:] = torch.rand((B, H, W, Lo.Cp - C), device=cuda, generator=g) * 40
    out = decode_topk(yp, _params(C, True, H, W), K=100)
    s, f, c = out["scores"], out["flat"], out["cls"]
    assert bool((s[:, :-1] >= s[:, 1:]).all())                                     # sorted by score
    tie = s[:, :-1] == s[:, 1:]
    assert bool((f[:, :-1][tie] < f[:, 1:][tie]).all())                            # ties by ascending flat index
    assert bool((c.long() == f % C).all())
    hm = yp[..., :C].reshape(B, -1)
    assert bool((hm.gather(1, f) == s).all())                                      # scores are the map values at flat
    mp = torch.nn.functional.max_pool2d(yp[..., :C].permute(0, 3, 1, 2), 3, 1, 1).permute(0, 2, 3, 1).reshape(B, -1)
    assert bool((mp.gather(1, f) == s).all())                                      # every winner is a 3x3 maximum
    # nothing better was missed: the K-th score bounds every other peak
    peaks = torch.where(hm == mp, hm, torch.zeros_like(hm))
    peaks.scatter_(1, f, 0.0)
    assert bool((peaks.max(dim=1).values <= s[:, -1]).all())
    out2 = decode_topk(yp, _params(C, True, H, W), K=100)                          # idempotent / deterministic
    assert all(torch.equal(out[k], out2[k]) for k in ("scores", "flat", "cls", "boxes"))
    ref = decode_np.decode_topk(Lo, yp[:3].cpu().numpy(), 100)
    _check_topk({k: v[:3] for k, v in out.items() if v is not None}, ref, 100)


def test_topk_multitask_wide_map(cuda):
    """BASELINE configs[4] layout: CenterNet channels [0:14] inside a 20-channel multitask tensor, W = 1536 (the 3x3
    neighbours of a pixel are too far apart to stay in the shared-memory ring: they are read from global memory), plus
    the semseg argmax of channels [14:19] of the same tensor."""
    from cvmhot import ops
    from cvmhot.layout import layout_from_params
    from cvmhot.common.utils.image import class_ids
    from oracle import image_np
    H, W, C, B = 24, 1536, 10, 3
    Lo = make_layout(H, W, C, "N")
    data = synth.make_batch(Lo, 9, B)
    rng = np.random.default_rng(3)
    wide = rng.normal(0, 1, (B, H, W, 20)).astype(np.float32)
    wide[..., :Lo.Cp] = data["y_pred"]
    ref = decode_np.decode_topk(Lo, data["y_pred"], 100)
    wide_d = torch.from_numpy(wide).to(cuda)
    out = ops.decode_topk(layout_from_params(_params(C, True, H, W)), wide_d[..., :Lo.Cp], K=100)
    _check_topk(out, ref, 100)
    ids = class_ids(wide_d, 14, 5).cpu().numpy()
    assert np.array_equal(ids, image_np.class_ids(wide[..., 14:19], 5))
    # the same two results from ONE pass over the wide tensor (cvm_decode_topk_semseg)
    wide[0, 3, 7, 14:19] = 0.25                        # all equal -> first index; a NaN is never "greater"
    wide[1, 5, 9, 14:19] = [0.1, np.nan, 0.3, 0.3, -1.0]
    wide_d = torch.from_numpy(wide).to(cuda)
    fused = ops.decode_topk(layout_from_params(_params(C, True, H, W)), wide_d[..., :Lo.Cp], K=100, semseg=(14, 5))
    _check_topk(fused, ref, 100)
    assert np.array_equal(fused["semseg_ids"].cpu().numpy(), class_ids(wide_d, 14, 5).cpu().numpy())
    assert fused["semseg_ids"][0, 3, 7] == 0 and fused["semseg_ids"][1, 5, 9] == 2


def test_topk_semseg_fused_generic_layout(cuda):
    """fused argmax on a layout without a compile-time instantiation (7 heatmap channels, 17-float pixels, odd sizes)."""
    from cvmhot import ops
    from cvmhot.layout import layout_from_params
    from cvmhot.common.utils.image import class_ids
    H, W, C, B = 37, 53, 7, 3
    Lo = make_layout(H, W, C, "N")
    data = synth.make_batch(Lo, 10, B)
    rng = np.random.default_rng(4)
    wide = rng.normal(0, 1, (B, H, W, 17)).astype(np.float32)
    wide[..., :Lo.Cp] = data["y_pred"]
    wide_d = torch.from_numpy(wide).to(cuda)
    ref = decode_np.decode_topk(Lo, data["y_pred"], 60)
    out = ops.decode_topk(layout_from_params(_params(C, True, H, W)), wide_d[..., :Lo.Cp], K=60, semseg=(Lo.Cp, 6))
    _check_topk(out, ref, 60)
    assert np.array_equal(out["semseg_ids"].cpu().numpy(), np.argmax(wide[..., Lo.Cp:Lo.Cp + 6], axis=-1).astype(np.uint8))


def test_topk_profile_r_class_from_logits(cuda):
    """Profile R (one objectness channel + class-logit field, the reference's default params): cls = first argmax of the
    class logits at the peak, like process_2d_output (post_processing.py:39-41), not flat % hm."""
    from cvmhot.models.centernet.post_processing import decode_topk
    H, W, nb, B = 40, 57, 6, 3
    Lo = make_layout(H, W, nb, "R")
    data = synth.make_batch(Lo, 12, B)
    yp = data["y_pred"]
    yp[0, 5, 5, 0] = 0.999
    yp[0, 5, 5, Lo.off_class:Lo.off_class + nb] = [0.5, 2.0, 2.0, -1.0, 0.0, 1.0]    # tie -> first
    ref = decode_np.decode_topk(Lo, yp, 50)
    out = decode_topk(torch.from_numpy(yp).to(cuda), _params(nb, False, H, W), K=50)
    _check_topk(out, ref, 50)
    assert ref["cls"].max() > 0 and int(out["cls"][0, 0]) == 1


@pytest.mark.parametrize("H,W,K_cls,B", [(40, 64, 40, 2), (33, 35, 3, 3)])
def test_topk_many_classes_and_unaligned(cuda, H, W, K_cls, B):
    """hm > 32 exercises the 64-bit channel masks; H*W*Cp not a multiple of 4 floats disables the bulk-copy engine (the
    loader warp copies with plain loads)."""
    from cvmhot.models.centernet.post_processing import decode_topk
    Lo = make_layout(H, W, K_cls, "N")
    data = synth.make_batch(Lo, 11, B)
    ref = decode_np.decode_topk(Lo, data["y_pred"], 100)
    out = decode_topk(torch.from_numpy(data["y_pred"]).to(cuda), _params(K_cls, True, H, W), K=100)
    _check_topk(out, ref, ref["scores"].shape[1])


def test_window9_vs_real_golden(cuda, golden_dir):
    from cvmhot.models.centernet import CenternetParams, process_2d_output
    from cvmhot.common.utils import Roi
    g = np.load(os.path.join(golden_dir, "decode_r.npz"))
    p = CenternetParams(int(g["nb_classes"]))
    roi = Roi(); roi.scale, roi.offset_left, roi.offset_top = [float(v) for v in g["roi"]]
    objs = process_2d_output(g["mask"], roi, p, float(g["min_conf"]))
    assert [o["cls_idx"] for o in objs] == list(g["cls"])
    assert np.array_equal(np.array([o["center"] for o in objs], np.float32), g["center"])
    assert np.array_equal(np.array([o["fullbox"] for o in objs], np.float32), g["fullbox"])


@pytest.mark.parametrize("H,W,nb", [(128, 384, 10), (9, 11, 3), (40, 57, 6)])
def test_window9_vs_oracle(cuda, H, W, nb):
    from cvmhot import ops
    from cvmhot.layout import layout_from_params
    Lo = make_layout(H, W, nb, "R")
    B = 3
    data = synth.make_batch(Lo, 6, B)
    yp = data["y_pred"]
    yp[..., 0] = np.maximum(yp[..., 0], (np.random.default_rng(1).uniform(0, 1, (B, H, W)) > 0.97) * 0.6).astype(np.float32)
    rois = [(0.25, -5, -3), (1.0, 0, 0), (0.5, 2, 9)]
    out = ops.decode_window9(layout_from_params(_params(nb, False, H, W)), torch.from_numpy(yp).to(cuda), 0.25,
                             ops.make_rois(rois, cuda), max_out=2000)
    for b in range(B):
        ref = decode_np.decode_window9(Lo, yp[b], rois[b], 0.25)
        n = int(out["counts"][b])
        assert n == len(ref)
        assert list(out["pix"][b, :n].cpu().numpy()) == [o["y"] * W + o["x"] for o in ref]      # scan order
        assert list(out["cls"][b, :n].cpu().numpy()) == [o["cls_idx"] for o in ref]
        if n:
            assert np.array_equal(out["centers"][b, :n].cpu().numpy(), np.array([o["center"] for o in ref], np.float32))
            assert np.array_equal(out["boxes"][b, :n].cpu().numpy(), np.array([o["fullbox"] for o in ref], np.float32))
            assert np.array_equal(out["scores"][b, :n].cpu().numpy(), np.array([o["score"] for o in ref], np.float32))


def test_topk_wrong_prediction_takes_the_exact_slow_path(cuda):
    """The scan kernel starts every image from a threshold PREDICTED from the image before it (and from the previous call,
    through the workspace).  A batch of high-scoring maps followed by low-scoring ones of the same shape makes every
    prediction too high: the merge kernel must notice (fewer than K published peaks reach the provisional threshold) and
    recompute those images exactly.  Results never depend on the prediction."""
    from cvmhot import ops
    from cvmhot.models.centernet.post_processing import decode_topk
    H, W, C, B = 64, 96, 10, 6
    Lo = make_layout(H, W, C, "N")
    rng = np.random.default_rng(21)
    hot = np.zeros((B, H, W, Lo.Cp), np.float32)
    hot[..., :C] = rng.uniform(0.5, 0.999, (B, H, W, C)).astype(np.float32)
    hot[..., C:] = rng.uniform(0, 50, (B, H, W, Lo.Cp - C)).astype(np.float32)
    cold = hot.copy()
    cold[..., :C] = rng.uniform(0.0, 0.2, (B, H, W, C)).astype(np.float32)
    cold[1, ..., :C] = 0.0                                   # no positive value at all
    cold[2, 5, 7, 3] = 0.9                                   # one lonely high peak
    p = _params(C, True, H, W)
    for _ in range(2):                                       # leaves a high prediction in the workspace
        out = decode_topk(torch.from_numpy(hot).to(cuda), p, K=100)
    _check_topk(out, decode_np.decode_topk(Lo, hot, 100), 100)
    n0 = ops.decode_fallback_count()
    out = decode_topk(torch.from_numpy(cold).to(cuda), p, K=100)
    assert ops.decode_fallback_count() > n0                  # the slow path ran ...
    _check_topk(out, decode_np.decode_topk(Lo, cold, 100), 100)   # ... and the result is exact
    mixed = np.concatenate([hot[:2], cold[:2], hot[2:4], cold[2:4]])          # inside one call: hot images predict for cold ones
    out = decode_topk(torch.from_numpy(mixed).to(cuda), p, K=100)
    _check_topk(out, decode_np.decode_topk(Lo, mixed, 100), 100)
    big = np.tile(mixed, (8, 1, 1, 1))                       # 64 images: several images per CTA, predictions chain inside a CTA
    out = decode_topk(torch.from_numpy(big).to(cuda), p, K=100)
    _check_topk(out, decode_np.decode_topk(Lo, big, 100), 100)
