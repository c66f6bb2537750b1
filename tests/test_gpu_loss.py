"""GPU parity: cvm_loss_fwd / cvm_loss_finalize / cvm_loss_bwd (through the C ABI) against the fp64 oracle.
Tolerance: 1e-5 relative (north_star)."""
import numpy as np
import pytest
import torch

import synth
from fixtures import reference_loss_fixture
from oracle import loss_np, render_np
from oracle.layout import make_layout

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _params(nb, per_class, H, W, track=False, l_shape=False, info3d=False):
    from cvmhot.models.centernet import CenternetParams
    from cvmhot.models.centertracker import CentertrackerParams
    p = (CentertrackerParams if track else CenternetParams)(nb, per_class)
    p.INPUT_HEIGHT, p.INPUT_WIDTH = H * p.R, W * p.R
    p.REGRESSION_FIELDS["l_shape"].active = l_shape
    p.REGRESSION_FIELDS["3d_info"].active = info3d
    return p


def _loss_cls(track):
    from cvmhot.models.centernet import CenternetLoss
    from cvmhot.models.centertracker import CentertrackerLoss
    return CentertrackerLoss if track else CenternetLoss


def _batch(Lo, config, B, track=False):
    data = synth.make_batch(Lo, config, B, track=track)
    yt = np.stack([render_np.render_image(Lo, data["boxes"][i], data["cls"][i], data["ignore"][i],
                                          data["track"][i] if track else None) for i in range(B)])
    return yt, data["y_pred"]


@pytest.mark.parametrize("profile,H,W,K,track,B", [("N", 128, 384, 10, False, 4), ("R", 128, 384, 10, False, 4),
                                                   ("N", 128, 384, 10, True, 3), ("R", 9, 11, 3, False, 2),
                                                   ("N", 37, 53, 7, False, 5), ("R", 128, 320, 6, True, 2)])
def test_loss_vs_oracle(cuda, profile, H, W, K, track, B):
    Lo = make_layout(H, W, K, profile, track=track)
    yt, yp = _batch(Lo, 3, B, track)
    loss = _loss_cls(track)(_params(K, profile == "N", H, W, track))
    yt_d, yp_d = torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)
    ref_total, ref_terms = loss_np.total_loss(Lo, yt, yp)
    got = loss(yt_d, yp_d)
    assert got.dim() == 0 and got.dtype == torch.float32
    assert float(got) == pytest.approx(ref_total, rel=RTOL)
    terms = loss.terms(yt_d, yp_d)
    assert float(terms["obj_focal"]) == pytest.approx(ref_terms[0], rel=RTOL)
    for f, v in zip(Lo.fields, ref_terms[1:]):
        assert float(terms[f.name]) == pytest.approx(v, rel=RTOL, abs=1e-12)
    # metric mode (train.py:62): unweighted focal; sub-terms accept y_true with or without the weights plane
    assert float(loss.obj_focal_loss(yt_d, yp_d)) == pytest.approx(loss_np.focal(Lo, yt, yp, use_weights=False), rel=RTOL)
    assert float(loss.obj_focal_loss(yt_d[..., :-1], yp_d, yt_d[..., -1])) == pytest.approx(ref_terms[0], rel=RTOL)
    assert float(loss.r_offset_loss(yt_d[..., :-1], yp_d)) == pytest.approx(ref_terms[1 + [f.name for f in Lo.fields].index("r_offset")], rel=RTOL)
    assert float(loss.fullbox_loss(yt_d, yp_d)) == pytest.approx(ref_terms[1 + [f.name for f in Lo.fields].index("fullbox")], rel=RTOL)
    # run-to-run bit reproducibility (no float atomics)
    assert float(loss(yt_d, yp_d)) == float(got)


def test_reference_fixture_all_fields(cuda):
    """The reference's own test input (loss_test.py:9-49) with every regression field active; KAT 0.27574 (App. C.1)."""
    from cvmhot.models.centernet import CenternetLoss
    nb, gt, perfect = reference_loss_fixture()
    Lo = make_layout(7, 7, 3, "R", l_shape=True, info3d=True)
    loss = CenternetLoss(_params(3, False, 7, 7, l_shape=True, info3d=True))
    gt_d = torch.from_numpy(gt).to(cuda)
    cases = {"perfect": perfect}
    p = perfect.copy(); p[0, 1, 1, 0] = 0.8; cases["peak08"] = p
    p = perfect.copy(); p[0, 2, 1, 0] = 1.0; cases["one_off"] = p
    p = perfect.copy(); p[0, 6, 1, 0] = 1.0; cases["wrong"] = p
    p = perfect.copy(); p[0, 1, 1, 5:20] += np.linspace(-1.5, 2.5, 15).astype(np.float32); cases["all_fields_off"] = p
    for name, pr in cases.items():
        ref_total, ref_terms = loss_np.total_loss(Lo, gt, pr)
        pr_d = torch.from_numpy(pr).to(cuda)
        assert float(loss(gt_d, pr_d)) == pytest.approx(ref_total, rel=RTOL), name
        t = loss.terms(gt_d, pr_d)
        for f, v in zip(Lo.fields, ref_terms[1:]):
            assert float(t[f.name]) == pytest.approx(v, rel=RTOL, abs=1e-9), (name, f.name)
    assert float(loss(gt_d, torch.from_numpy(cases["perfect"]).to(cuda))) == pytest.approx(0.27574, abs=2e-5)
    assert float(loss.calc_loss(gt_d[..., :-1], gt_d[..., 4:6], torch.from_numpy(cases["all_fields_off"]).to(cuda)[..., 4:6], "mae")) == \
        pytest.approx(loss_np.total_loss(Lo, gt, cases["all_fields_off"])[1][2], rel=RTOL)


def test_tracker_fixture(cuda):
    from cvmhot.models.centertracker import CentertrackerLoss
    nb, gt, perfect = reference_loss_fixture(track=True)
    Lo = make_layout(7, 7, 3, "R", track=True, l_shape=True, info3d=True)
    loss = CentertrackerLoss(_params(3, False, 7, 7, track=True, l_shape=True, info3d=True))
    p = perfect.copy(); p[0, 1, 1, Lo.off_track:Lo.off_track + 2] = [0.0, 3.0]
    gt_d = torch.from_numpy(gt).to(cuda)
    assert float(loss(gt_d, torch.from_numpy(p).to(cuda))) == pytest.approx(loss_np.total_loss(Lo, gt, p)[0], rel=RTOL)
    assert float(loss.track_offset_loss(gt_d, torch.from_numpy(p).to(cuda))) == pytest.approx(2.0, rel=RTOL)


def test_no_objects_and_empty_batch(cuda):
    from cvmhot.models.centernet import CenternetLoss
    Lo = make_layout(16, 24, 4, "N")
    rng = np.random.default_rng(1)
    yt = np.zeros((2, 16, 24, Lo.Ct), np.float32); yt[..., -1] = 1
    yt[..., :4] = rng.uniform(0, 0.9, (2, 16, 24, 4))
    yp = rng.uniform(0.0, 1.0, (2, 16, 24, Lo.Cp)).astype(np.float32)
    yp[0, 0, 0, 0], yp[0, 0, 1, 0] = 0.0, 1.0         # clip edges
    loss = CenternetLoss(_params(4, True, 16, 24))
    got = float(loss(torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)))
    assert got == pytest.approx(loss_np.total_loss(Lo, yt, yp)[0], rel=RTOL)


def test_multitask_slice_strides(cuda):
    """CenterNet slice inside wider multitask tensors (multitask/loss.py:21-24,44-47) — read in place, no copy."""
    from cvmhot.models.multitask import MultitaskParams, MultitaskLoss
    mp = MultitaskParams(10, per_class_heatmap=True)
    H, W = 32, 48
    mp.cn_params.INPUT_HEIGHT, mp.cn_params.INPUT_WIDTH = H * 2, W * 2
    Lo = make_layout(H, W, 10, "N")
    yt, yp = _batch(Lo, 5, 3)
    rng = np.random.default_rng(2)
    yt_w = np.concatenate([yt, rng.uniform(0, 1, (3, H, W, 7)).astype(np.float32)], axis=-1)      # + semseg(5+1) + depth
    yp_w = np.concatenate([yp, rng.uniform(0, 1, (3, H, W, 6)).astype(np.float32)], axis=-1)
    assert yt_w.shape[-1] == 22 and yp_w.shape[-1] == 20                                        # config 5 strides
    ml = MultitaskLoss(mp)
    got = float(ml.calc_centernet(torch.from_numpy(yt_w).to(cuda), torch.from_numpy(yp_w).to(cuda)))
    assert got == pytest.approx(loss_np.total_loss(Lo, yt, yp)[0], rel=RTOL)
    ids = ml.semseg_class_ids(torch.from_numpy(yp_w).to(cuda)).cpu().numpy()
    assert np.array_equal(ids, np.argmax(yp_w[..., 14:19], axis=-1).astype(np.uint8))


def _torch_loss64(Lo, yt, yp):
    """differentiable fp64 torch restatement (CPU) used only to check the hand-written backward."""
    w = yt[..., -1:]
    Y, Yh = yt[..., :Lo.hm], yp[..., :Lo.hm]
    pos, neg = (Y == 1.0).double(), (Y < 1.0).double()
    pl = -pos * (1 - Yh) ** Lo.focal_a * torch.log(torch.clamp(Yh, 0.01, 0.99))
    nl = -neg * (1 - Y) ** Lo.focal_b * Yh ** Lo.focal_a * torch.log(torch.clamp(1 - Yh, 0.01, 0.99))
    n = pos.sum()
    total = ((pl * w).sum() + (nl * w).sum()) / n if n > 0 else (nl * w).sum()
    pm = pos.max(dim=-1, keepdim=True).values
    nobj = pm.sum()
    for f in Lo.fields:
        t, p = yt[..., f.off:f.off + f.size], yp[..., f.off:f.off + f.size]
        if f.kind == 0:
            m = pm * (t - p) ** 2
        elif f.kind == 1:
            m = pm * (t - p)
        elif f.kind == 2:
            m = pm * ((t - p) / torch.clamp(t.abs(), min=1.0))
        else:
            m = -(t * torch.log_softmax(p, dim=-1)).sum(-1) * pm[..., 0]
        v = m.abs().sum()
        v = v / nobj if nobj > 0 else v
        if f.post == 1:
            v = torch.sqrt(1.0 - 0.99 * torch.cos(2.0 * v)) + (v * v * 0.05).abs() - 0.0999
        total = total + v * f.weight
    return total


@pytest.mark.parametrize("profile,track,allf,K,H,W", [
    ("N", False, False, 5, 24, 40), ("R", True, False, 5, 24, 40), ("R", False, True, 5, 24, 40),
    # the layouts the fast backward kernel is instantiated for (10 classes: 15/14 and tracker 17/16 channels); 25x41x3
    # pixels = 19 full 160-pixel spans for the fast kernel + a ragged tail for the generic one
    ("N", False, False, 10, 25, 41), ("N", True, False, 10, 25, 41), ("N", False, False, 10, 32, 40)])
def test_backward_vs_autograd(cuda, profile, track, allf, K, H, W):
    B = 3
    Lo = make_layout(H, W, K, profile, track=track, l_shape=allf, info3d=allf)
    yt, yp = _batch(Lo, 9, B, track)
    if allf:   # give the extra fields targets at the peaks
        rng = np.random.default_rng(3)
        pk = (yt[..., :Lo.hm] == 1).any(-1)
        for f in Lo.fields:
            if f.name in ("l_shape", "radial_dist", "orientation", "obj_dims"):
                yt[pk, f.off:f.off + f.size] = rng.normal(0, 2, (int(pk.sum()), f.size)).astype(np.float32)
    yp[0, 0, 0, 0], yp[0, 0, 1, 0] = 0.005, 0.995     # outside the clip range: zero log-gradient
    loss = _loss_cls(track)(_params(K, profile == "N", H, W, track, allf, allf))
    yp_d = torch.from_numpy(yp).to(cuda).requires_grad_(True)
    out = loss(torch.from_numpy(yt).to(cuda), yp_d)
    (out * 3.0).backward()
    yp64 = torch.from_numpy(yp).double().requires_grad_(True)
    ref = _torch_loss64(Lo, torch.from_numpy(yt).double(), yp64)
    (ref * 3.0).backward()
    assert float(out.detach()) == pytest.approx(float(ref.detach()), rel=RTOL)
    g, gr = yp_d.grad.cpu().double(), yp64.grad
    scale = gr.abs().max()
    # fp64 autograd of the fp64 restatement = the analytic gradient; 1e-5 relative (north_star), with an absolute floor of
    # 1e-6 of the largest entry for the elements that are differences of nearly equal terms
    assert torch.allclose(g, gr, rtol=1e-5, atol=float(scale) * 1e-6)


def test_backward_full_size_fast_equals_generic(cuda, monkeypatch):
    """BASELINE configs[1] shape (B = 16): the pipelined backward kernel (compile-time layout) against the generic one on the
    same inputs, and two size-independent properties: the gradient is linear in the upstream gradient and it is zero
    wherever y_pred has no loss term (regression channels of non-peak pixels)."""
    import bench
    from cvmhot import ops
    from cvmhot.layout import layout_from_params
    from cvmhot.models.centernet import CenternetParams
    H, W, K, B = 128, 384, 10, 16
    p = CenternetParams(K, True)
    p.INPUT_HEIGHT, p.INPUT_WIDTH = H * p.R, W * p.R
    L = layout_from_params(p)
    Lo = make_layout(H, W, K, "N")
    boxes, cls, ign = bench.gen_objects(3, B)
    yt = np.stack([render_np.render_image(Lo, boxes[b], cls[b], ign[b]) for b in range(B)])
    g = torch.Generator(device=cuda).manual_seed(5)
    yt_d = torch.from_numpy(yt).to(cuda)
    yp_d = torch.empty((B, H, W, Lo.Cp), device=cuda)
    yp_d[..., :K] = torch.sigmoid(torch.randn((B, H, W, K), device=cuda, generator=g) * 2.5 - 2.0)   # some outside [.01,.99]
    yp_d[..., K:] = torch.rand((B, H, W, Lo.Cp - K), device=cuda, generator=g) * 40
    part = ops.loss_partials(L, yt_d, yp_d, True).clone()
    fast = ops.loss_backward(L, yt_d, yp_d, part).clone()
    generic = ops.loss_backward(L, yt_d, yp_d, part, generic=True).clone()
    assert torch.isfinite(fast).all()
    scale = float(generic.abs().max())
    assert torch.allclose(fast, generic, rtol=1e-5, atol=scale * 1e-7)
    up = torch.tensor(2.0, device=cuda)
    twice = ops.loss_backward(L, yt_d, yp_d, part, upstream=up)
    assert torch.allclose(twice, 2.0 * fast, rtol=1e-6, atol=0)
    peak = (yt_d[..., :K] == 1.0).any(-1)
    assert int(peak.sum()) > 0
    assert float(fast[..., K:][~peak].abs().max()) == 0.0
    assert float(fast[..., K:][peak].abs().max()) > 0.0


def test_loss_properties_full_size(cuda):
    """BASELINE configs[1] shape (B sampled down to 32): the partials vector is additive over batch shards (this is what the
    multi-GPU all-reduce relies on), bit-reproducible run to run, and the total agrees with the fp64 oracle."""
    import bench
    from cvmhot import ops
    from cvmhot.layout import layout_from_params
    from cvmhot.models.centernet import CenternetParams
    H, W, K, B = 128, 384, 10, 32
    p = CenternetParams(K, True)
    p.INPUT_HEIGHT, p.INPUT_WIDTH = H * p.R, W * p.R
    L = layout_from_params(p)
    Lo = make_layout(H, W, K, "N")
    boxes, cls, ign = bench.gen_objects(0, B)
    yt = np.stack([render_np.render_image(Lo, boxes[b], cls[b], ign[b]) for b in range(B)])
    rng = np.random.default_rng(21)
    yp = np.zeros((B, H, W, Lo.Cp), np.float32)
    yp[..., :K] = 1.0 / (1.0 + np.exp(-rng.normal(-4.0, 1.5, (B, H, W, K))))
    yp[..., K:] = rng.uniform(0, 60, (B, H, W, Lo.Cp - K))
    yt_d, yp_d = torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)
    whole = ops.loss_partials(L, yt_d, yp_d, True).clone()
    again = ops.loss_partials(L, yt_d, yp_d, True).clone()
    assert torch.equal(whole, again)                                               # fixed-order reductions: reproducible
    parts = sum(ops.loss_partials(L, yt_d[a:b], yp_d[a:b], True).clone() for a, b in ((0, 5), (5, 20), (20, 32)))
    # additive over shards up to the rounding of the short fp32 chains inside a thread (everything above them is fp64)
    assert torch.allclose(parts, whole, rtol=2e-6, atol=0)
    total = float(ops.loss_finalize(L, whole)[0])
    assert total == pytest.approx(loss_np.total_loss(Lo, yt, yp)[0], rel=RTOL)


def test_targets_above_one_are_masked(cuda):
    """Y > 1 belongs to neither the positive nor the negative mask (loss.py:35-36).  The render never produces it, but the
    kernel must still honour it (its packed fast path redoes such pixels channel by channel)."""
    from cvmhot import ops
    from cvmhot.layout import layout_from_params
    from cvmhot.models.centernet import CenternetParams
    H, W, K, B = 16, 24, 10, 2
    p = CenternetParams(K, True)
    p.INPUT_HEIGHT, p.INPUT_WIDTH = H * p.R, W * p.R
    Lo = make_layout(H, W, K, "N")
    data = synth.make_batch(Lo, 5, B)
    yt = np.stack([render_np.render_image(Lo, data["boxes"][b], data["cls"][b], data["ignore"][b]) for b in range(B)])
    yt[0, 3, 4, 2] = 1.5
    yt[1, 7, 7, 0] = 2.0
    yt[1, 7, 7, 9] = 1.0000001
    out = ops.loss_finalize(layout_from_params(p), ops.loss_partials(layout_from_params(p), torch.from_numpy(yt).to(cuda),
                                                                     torch.from_numpy(data["y_pred"]).to(cuda), True))
    assert float(out[0]) == pytest.approx(loss_np.total_loss(Lo, yt, data["y_pred"])[0], rel=RTOL)


# ---- against values produced by EXECUTING the reference's loss source (tests/golden/loss_*.npz) -----------------------
import loss_golden  # noqa: E402

_GOLD = {c[0]: c for c in loss_golden.cases()}


@pytest.mark.parametrize("name", loss_golden.case_ids())
def test_cuda_loss_vs_executed_reference(cuda, name):
    """CUDA path (through the drop-in mirror and the C ABI) against the unmodified reference loss.py / centertracker
    loss.py run over oracle/tf_shim.py (make_golden.py): total, weighted focal, metric-mode focal and every field term."""
    _, kw, yt, yp, vals = _GOLD[name]
    H, W = yt.shape[1], yt.shape[2]
    loss = _loss_cls(kw.get("track", False))(loss_golden.product_params(kw, H, W))
    yt_d, yp_d = torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)
    assert float(loss(yt_d, yp_d)) == pytest.approx(vals["total"], rel=RTOL, abs=1e-7)
    assert float(loss.obj_focal_loss(yt_d[..., :-1], yp_d, yt_d[..., -1])) == pytest.approx(vals["focal_weighted"], rel=RTOL, abs=1e-7)
    assert float(loss.obj_focal_loss(yt_d, yp_d)) == pytest.approx(vals["focal_metric"], rel=RTOL, abs=1e-7)
    for term in vals:
        if term in ("total", "focal_weighted", "focal_metric"):
            continue
        got = float(getattr(loss, term + "_loss")(yt_d[..., :-1], yp_d))
        assert got == pytest.approx(vals[term], rel=RTOL, abs=1e-7), term


def test_cuda_multitask_vs_executed_reference(cuda, golden_dir):
    """MultitaskLoss.calc_centernet on the wide tensors (strided in-place read) against the reference's own
    MultitaskLoss.calc_centernet (models/multitask/loss.py:44-47) run over the shim."""
    from cvmhot.models.multitask import MultitaskParams, MultitaskLoss
    z = np.load(golden_dir + "/loss_multitask.npz")
    yt, yp = z["y_true"], z["y_pred"].astype(np.float32)
    mp = MultitaskParams(int(z["nb_classes"]))
    mp.cn_params.INPUT_HEIGHT, mp.cn_params.INPUT_WIDTH = yt.shape[1] * 2, yt.shape[2] * 2
    ml = MultitaskLoss(mp)
    assert ml.cn_offset["y_true"] == list(z["cn_offset_true"]) and ml.cn_offset["y_pred"] == list(z["cn_offset_pred"])
    assert ml.semseg_offset["y_pred"] == list(z["semseg_offset_pred"]) and ml.depth_offset["y_pred"] == list(z["depth_offset_pred"])
    got = float(ml.calc_centernet(torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)))
    assert got == pytest.approx(float(z["total"]), rel=RTOL)


def test_metric_terms_share_one_pass(cuda, monkeypatch):
    """The reference's metrics list (train.py:62) calls every sub-term on the same batch: the drop-in runs ONE fused pass for
    all of them, a changed tensor (in place or a new one) triggers a new pass, and the single-device path is one launch
    (cvm_loss_fwd_total) whose result equals partials + finalize."""
    from cvmhot import ops
    Lo = make_layout(24, 40, 5, "R")
    yt, yp = _batch(Lo, 13, 2)
    loss = _loss_cls(False)(_params(5, False, 24, 40))
    yt_d, yp_d = torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)
    calls = []
    real = ops.loss_total
    monkeypatch.setattr(ops, "loss_total", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    vals = [float(f(yt_d, yp_d)) for f in (loss.obj_focal_loss, loss.class_loss, loss.r_offset_loss, loss.fullbox_loss)]
    vals2 = [float(f(yt_d[..., :-1], yp_d)) for f in (loss.class_loss, loss.r_offset_loss, loss.fullbox_loss)]
    assert len(calls) == 1 and vals[1:] == vals2
    ref_total, ref_terms = loss_np.total_loss(Lo, yt, yp)
    assert vals[1] == pytest.approx(ref_terms[1], rel=RTOL) and vals[3] == pytest.approx(ref_terms[3], rel=RTOL)
    yp_d[0, 3, 3, Lo.off_box] += 5.0                      # in-place change: the cached vector must not be served
    yp[0, 3, 3, Lo.off_box] += 5.0
    pk = (yt[..., 0] == 1.0)
    yp_d[torch.from_numpy(pk).to(cuda)] += 1.0
    yp[pk] += 1.0
    assert float(loss.fullbox_loss(yt_d, yp_d)) == pytest.approx(loss_np.total_loss(Lo, yt, yp)[1][3], rel=RTOL)
    assert len(calls) == 2
    yp2_d = yp_d.clone()                                  # a new tensor: new pass
    loss.fullbox_loss(yt_d, yp2_d)
    assert len(calls) == 3
    # one launch == two launches
    L = loss._layout(yt_d)
    out1, part1 = real(L, yt_d, yp_d, True)
    part2 = ops.loss_partials(L, yt_d, yp_d, True)
    out2 = ops.loss_finalize(L, part2)
    assert torch.equal(part1, part2) and torch.equal(out1[:2 + len(Lo.fields)], out2[:2 + len(Lo.fields)])
