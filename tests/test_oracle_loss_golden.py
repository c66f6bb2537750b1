"""CPU: the fp64 loss restatement (oracle/loss_np.py) against values produced by EXECUTING the reference's loss source
(tests/golden/loss_*.npz, see tests/golden/make_golden.py and oracle/tf_shim.py).  Tolerance 1e-5 relative
(north_star); the reference side ran in fp32 like TF does."""
import numpy as np
import pytest

import loss_golden
from oracle import loss_np, ref_import

RTOL = 1e-5
_CASES = {c[0]: c for c in loss_golden.cases()}


@pytest.mark.parametrize("name", loss_golden.case_ids())
def test_oracle_matches_executed_reference(name):
    _, kw, yt, yp, vals = _CASES[name]
    L = loss_golden.oracle_layout(kw, yt.shape[1], yt.shape[2])
    assert (L.Ct, L.Cp) == (yt.shape[-1], yp.shape[-1])
    total, terms = loss_np.total_loss(L, yt, yp)
    assert total == pytest.approx(vals["total"], rel=RTOL, abs=1e-7)
    assert terms[0] == pytest.approx(vals["focal_weighted"], rel=RTOL, abs=1e-7)
    assert loss_np.focal(L, yt, yp, use_weights=False) == pytest.approx(vals["focal_metric"], rel=RTOL, abs=1e-7)
    seen = 0
    for f, v in zip(L.fields, terms[1:]):
        assert v == pytest.approx(vals[f.name], rel=RTOL, abs=1e-7), f.name
        seen += 1
    assert seen == len(vals) - 3          # every term the reference produced was compared


def test_multitask_slice_matches_executed_reference(golden_dir):
    z = np.load(golden_dir + "/loss_multitask.npz")
    yt, yp = z["y_true"], z["y_pred"].astype(np.float32)
    a, b = z["cn_offset_true"]
    c, d = z["cn_offset_pred"]
    from oracle.layout import make_layout
    L = make_layout(yt.shape[1], yt.shape[2], int(z["nb_classes"]), "R")
    assert (b - a, d - c) == (L.Ct, L.Cp)
    total, _ = loss_np.total_loss(L, yt[..., a:b], yp[..., c:d])
    assert total == pytest.approx(float(z["total"]), rel=RTOL)


@pytest.mark.skipif(not ref_import.available(), reason="reference not mounted (GPU box)")
def test_goldens_regenerate_from_live_reference(golden_dir):
    """Re-run the unmodified reference loss over the shim on the stored inputs: the committed values are reproducible."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", golden_dir + "/make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    r = ref_import.load()
    import inspect
    src = inspect.getsourcefile(r["CenternetLoss"])
    assert src.startswith(ref_import.REFERENCE_ROOT)          # the class that ran is the reference's own file
    for name in ("profile_r_128x384", "tracker_n_64x96", "all_fields_r_32x48", "trk_track_off", "no_objects_r_16x24"):
        _, kw, yt, yp, vals = _CASES[name]
        hm = kw["nb"] if kw["profile"] == "N" else 1
        loss, _ = mg.ref_loss_object(r, kw["nb"], hm=hm, track=kw.get("track", False), l_shape=kw.get("l_shape", False),
                                     info3d=kw.get("info3d", False))
        got = mg.ref_loss_values(loss, yt, yp)
        for k, v in vals.items():
            assert got[k] == pytest.approx(v, rel=1e-6, abs=1e-9), (name, k)
