"""Seeded shape fuzzing of the three kernels against the oracle: odd sizes, 1..40 classes, tracker / Profile-R layouts,
K around the element count, single images and ragged batches.  The cases are small (the oracle is NumPy) but walk every
code path of the kernels: ring / global neighbours, one or two granules per step, bulk and plain copies, partial granules
and chunks, empty images."""
import numpy as np
import pytest
import torch

import synth
from oracle import decode_np, loss_np, render_np
from oracle.layout import make_layout

pytestmark = pytest.mark.gpu


def _params(nb, per_class, H, W, track=False):
    from cvmhot.models.centernet import CenternetParams
    from cvmhot.models.centertracker import CentertrackerParams
    p = (CentertrackerParams if track else CenternetParams)(nb, per_class)
    p.INPUT_HEIGHT, p.INPUT_WIDTH = H * p.R, W * p.R
    return p


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        H, W = int(rng.integers(6, 70)), int(rng.integers(6, 90))
        if rng.random() < 0.25:
            W = int(rng.integers(400, 1300))      # wider than a granule: neighbours far apart
            H = int(rng.integers(6, 14))
        out.append(dict(H=H, W=W, K_cls=int(rng.choice([1, 2, 3, 5, 10, 17, 33, 40])), B=int(rng.integers(1, 6)),
                        track=bool(rng.random() < 0.3), K=int(rng.choice([1, 7, 50, 100, 333])),
                        cfg=int(rng.integers(20, 1000))))
    return out


@pytest.mark.parametrize("c", _cases(28, 101), ids=lambda c: f"{c['H']}x{c['W']}x{c['K_cls']}-B{c['B']}-K{c['K']}-t{int(c['track'])}")
def test_decode_fuzz(cuda, c):
    from cvmhot.models.centernet.post_processing import decode_topk
    Lo = make_layout(c["H"], c["W"], c["K_cls"], "N", track=c["track"])
    data = synth.make_batch(Lo, c["cfg"], c["B"], track=c["track"])
    yp = data["y_pred"]
    if c["cfg"] % 3 == 0:                                   # sparse maps: long score-0 tails
        yp[..., :Lo.hm] *= (np.random.default_rng(c["cfg"]).random(yp[..., :Lo.hm].shape) < 0.02)
    ref = decode_np.decode_topk(Lo, yp, c["K"])
    out = decode_topk(torch.from_numpy(yp).to(cuda), _params(c["K_cls"], True, c["H"], c["W"], c["track"]), K=c["K"])
    K_eff = ref["scores"].shape[1]
    for k in ("scores", "cls", "flat", "centers", "boxes"):
        assert np.array_equal(out[k].cpu().numpy()[:, :K_eff], ref[k]), k
    if c["track"]:
        assert np.array_equal(out["track"].cpu().numpy()[:, :K_eff], ref["track"])


@pytest.mark.parametrize("c", _cases(14, 202), ids=lambda c: f"{c['H']}x{c['W']}x{c['K_cls']}-B{c['B']}-t{int(c['track'])}")
def test_render_and_loss_fuzz(cuda, c):
    from cvmhot.layout import layout_from_params
    from cvmhot.models.centernet.processor import ProcessImages, pack_objects, pack_boxes
    from cvmhot.models.centernet import CenternetLoss
    from cvmhot.models.centertracker import CentertrackerLoss, CenterTrackerProcess
    profile = "R" if c["cfg"] % 2 else "N"
    p = _params(c["K_cls"], profile == "N", c["H"], c["W"], c["track"])
    Lo = make_layout(c["H"], c["W"], c["K_cls"], profile, track=c["track"])
    L = layout_from_params(p)
    data = synth.make_batch(Lo, c["cfg"], c["B"], track=c["track"])
    proc = CenterTrackerProcess(p) if c["track"] else ProcessImages(p)
    rec, offs = pack_objects(data["boxes"], data["cls"], data["track"])
    y = proc.render_packed(L, rec, offs, *pack_boxes(data["ignore"]))
    ref = np.stack([render_np.render_image(Lo, data["boxes"][b], data["cls"][b], data["ignore"][b],
                                           data["track"][b] if c["track"] else None) for b in range(c["B"])])
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-5, atol=1e-30)
    loss = (CentertrackerLoss if c["track"] else CenternetLoss)(p)
    got = float(loss(y, torch.from_numpy(data["y_pred"]).to(cuda)))
    assert got == pytest.approx(loss_np.total_loss(Lo, ref, data["y_pred"])[0], rel=1e-5)
