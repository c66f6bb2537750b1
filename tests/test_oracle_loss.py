"""CPU: the fp64 loss restatement.  The reference's loss tests carry no numeric expectations (inequalities only, some
stale), so this pins the restatement by (a) hand-derived closed forms, (b) the reference's fixture
with its still-valid inequalities, (c) an independently written torch fp32 twin."""
import math

import numpy as np
import pytest

from oracle import loss_np
from oracle.layout import make_layout
from fixtures import reference_loss_fixture
import synth


def _layout_all(track=False):
    return make_layout(7, 7, 3, "R", track=track, l_shape=True, info3d=True)


def test_reference_fixture_known_answers():
    nb, gt, pred = reference_loss_fixture()
    L = _layout_all()
    assert L.Cp == 20 and gt.shape[-1] == 21
    total, terms = loss_np.total_loss(L, gt, pred)
    # focal: the positive is predicted 1.0 -> (1-1)^2 = 0; negatives are predicted 0 -> 0^2 = 0
    assert terms[0] == 0.0
    # class CE on logits [0,1,0] vs one-hot[1]: -log softmax = log(1 + 2/e) ; all other fields are exact
    ce = math.log(1.0 + 2.0 / math.e)
    assert terms[1] == pytest.approx(ce, rel=1e-12)
    assert terms[2] == 0.0 and terms[3] == 0.0 and terms[4] == 0.0 and terms[5] == 0.0 and terms[7] == 0.0
    orient0 = math.sqrt(1.0 - 0.99) - 0.0999          # orientation_loss(0), loss.py:98
    assert terms[6] == pytest.approx(orient0, rel=1e-9)
    assert total == pytest.approx(0.5 * ce + 0.2 * orient0, rel=1e-12)
    assert total == pytest.approx(0.27574, abs=2e-5)  # value derived in SURVEY.md App. C.1


def test_reference_fixture_inequalities():
    """loss_test.py:62-68,84-145: the inequalities that hold for the current loss definition."""
    nb, gt, perfect = reference_loss_fixture()
    L = _layout_all()
    loss = lambda p: loss_np.total_loss(L, gt, p)[0]
    p = perfect.copy(); p[0, 2, 1, 0] = 1.0
    one_off = loss(p)
    p = perfect.copy(); p[0, 6, 1, 0] = 1.0
    wrong = loss(p)
    assert one_off < wrong                              # loss_test.py:68
    assert one_off == pytest.approx(0.28311, abs=2e-5) and wrong == pytest.approx(4.88091, abs=2e-5)   # App. C.1
    p = perfect.copy(); p[0, 1, 1, 0] = 0.8
    assert loss(p) == pytest.approx(0.28467, abs=2e-5)  # App. C.1
    bx = L.off_box
    p = perfect.copy(); p[0, 1, 1, bx] = 2.2 + 10
    small = loss(p)
    p[0, 1, 1, bx] = 2.2 - 30
    large = loss(p)
    assert large > small > 0                            # loss_test.py:84-91
    f = {x.name: x for x in L.fields}
    p = perfect.copy(); p[0, 1, 1, f["l_shape"].off:f["l_shape"].off + 7] += 1
    assert loss_np.field_loss(L, f["l_shape"], gt, p) > 1.0          # :118
    p = perfect.copy(); p[0, 1, 1, f["radial_dist"].off] += 2.5
    assert loss_np.field_loss(L, f["radial_dist"], gt, p) > 0.1      # :126
    p = perfect.copy(); p[0, 1, 1, f["orientation"].off] += 1.5
    assert loss_np.field_loss(L, f["orientation"], gt, p) > 0.1      # :134
    p = perfect.copy(); p[0, 1, 1, f["obj_dims"].off:f["obj_dims"].off + 3] += 0.5
    assert loss_np.field_loss(L, f["obj_dims"], gt, p) > 0.1         # :145


def test_tracker_fixture():
    nb, gt, perfect = reference_loss_fixture(track=True)
    L = _layout_all(track=True)
    assert L.Cp == 22
    base = loss_np.total_loss(L, gt, perfect)[0]
    p = perfect.copy(); p[0, 1, 1, L.off_track:L.off_track + 2] = [0.0, 3.0]
    some = loss_np.total_loss(L, gt, p)[0]
    assert some - base == pytest.approx(0.1 * 2.0, rel=1e-9)        # mse: (1-0)^2 + (2-3)^2 = 2, weight 0.1
    assert some - base > 0.1                                         # centertracker/loss_test.py:31 (relative to baseline)


def test_no_objects_branch():
    """tf.cond(n > 0, ...) false branch (loss.py:59,130): sums are returned un-normalised."""
    L = make_layout(8, 8, 4, "N")
    rng = np.random.default_rng(0)
    yt = np.zeros((2, 8, 8, L.Ct), np.float32); yt[..., -1] = 1
    yt[..., :4] = rng.uniform(0, 0.9, (2, 8, 8, 4))
    yp = rng.uniform(0.02, 0.9, (2, 8, 8, L.Cp)).astype(np.float32)
    part = loss_np.partials(L, yt, yp)
    assert part[2] == 0 and part[3] == 0 and part[0] == 0
    total, terms = loss_np.finalize(L, part)
    assert total == pytest.approx(part[1]) and terms[1] == 0 and terms[2] == 0


@pytest.mark.parametrize("profile,track", [("N", False), ("R", False), ("N", True)])
def test_fp32_twin_agrees(profile, track):
    from oracle import render_np
    L = make_layout(32, 48, 5, profile, track=track)
    data = synth.make_batch(L, 7, 3, n_obj=6, track=track)
    yt = np.stack([render_np.render_image(L, data["boxes"][i], data["cls"][i], data["ignore"][i],
                                          data["track"][i] if track else None) for i in range(3)])
    a = loss_np.total_loss(L, yt, data["y_pred"])[0]
    b = loss_np.loss_torch32(L, yt, data["y_pred"])
    assert a == pytest.approx(b, rel=1e-5)
    # sharding invariance: partials add up (this is what the all-reduce relies on)
    p = loss_np.partials(L, yt[:1], data["y_pred"][:1]) + loss_np.partials(L, yt[1:], data["y_pred"][1:])
    assert loss_np.finalize(L, p)[0] == pytest.approx(a, rel=1e-12)
