"""cvmhot/keras_shim.py (the tf.keras.losses.Loss front of the CUDA loss) driven through the TensorFlow stand-in of
oracle/tf_shim.py: TensorFlow is not in the image, so this exercises the shim's glue - DLPack hand-over, py_function
bodies, the custom-gradient pair, the metric methods - with torch tensors standing in for EagerTensors.  It does not prove
anything about real TensorFlow's graph tracing."""
import importlib
import sys

import pytest
import torch

import loss_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-5
_GOLD = {c[0]: c for c in loss_golden.cases()}


@pytest.fixture()
def shim(cuda):
    from oracle import tf_shim
    had = sys.modules.get("tensorflow")
    tf_shim.install()
    sys.modules.pop("cvmhot.keras_shim", None)
    mod = importlib.import_module("cvmhot.keras_shim")
    yield mod
    sys.modules.pop("cvmhot.keras_shim", None)
    if had is None:
        for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
            sys.modules.pop(k)


@pytest.mark.parametrize("name", ["profile_r_128x384", "tracker_n_64x96", "all_fields_r_32x48"])
def test_keras_shim_call_metrics_and_gradient(shim, cuda, name):
    from cvmhot.models.centernet.loss import CenternetLoss as TorchLoss
    from cvmhot.models.centertracker.loss import CentertrackerLoss as TorchTrackerLoss
    _, kw, yt, yp, vals = _GOLD[name]
    track = kw.get("track", False)
    params = loss_golden.product_params(kw, yt.shape[1], yt.shape[2])
    yt_d, yp_d = torch.from_numpy(yt).to(cuda), torch.from_numpy(yp).to(cuda)
    loss = (shim.CentertrackerLoss if track else shim.CenternetLoss)(params)
    assert isinstance(loss, sys.modules["tensorflow"].keras.losses.Loss)
    total = loss(yt_d, yp_d)                                   # keras.losses.Loss.__call__ -> call()
    assert float(total) == pytest.approx(vals["total"], rel=RTOL, abs=1e-7)        # the EXECUTED reference's value
    # the metric methods of train.py:62 (they get y_true with the weights plane attached)
    assert float(loss.obj_focal_loss(yt_d, yp_d)) == pytest.approx(vals["focal_metric"], rel=RTOL, abs=1e-7)
    for term in vals:
        if term in ("total", "focal_weighted", "focal_metric"):
            continue
        got = float(getattr(loss, term + "_loss")(yt_d, yp_d))
        assert got == pytest.approx(vals[term], rel=RTOL, abs=1e-7), term
    with pytest.raises(NotImplementedError):
        loss.obj_focal_loss(yt_d[..., :-1], yp_d, yt_d[..., -1])
    # gradient: the function tf.custom_gradient would hand to the tape equals the torch mirror's autograd
    g = total._grad_fn_for_tape(torch.ones((), device=cuda))
    ref = (TorchTrackerLoss if track else TorchLoss)(params)
    ypr = yp_d.clone().requires_grad_(True)
    ref.call(yt_d, ypr).backward()
    assert g.shape == yp_d.shape
    assert torch.equal(g, ypr.grad)
    g2 = total._grad_fn_for_tape(torch.full((), 0.5, device=cuda))
    torch.testing.assert_close(g2, 0.5 * ypr.grad, rtol=1e-6, atol=0)
