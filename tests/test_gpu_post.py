"""GPU parity: cvm_semseg_argmax against the REAL to_3channel outputs (golden) and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import image_np

pytestmark = pytest.mark.gpu


def test_to_3channel_vs_real_golden(cuda, golden_dir):
    from cvmhot.common.utils import to_3channel
    g = np.load(os.path.join(golden_dir, "to3.npz"))
    cols = [tuple(int(v) for v in c) for c in g["colours"]]
    items = [(n, c) for n, c in zip("abcde", cols)]
    for thr in (None, 0.3):
        for uw in (False, True):
            for sm in (True, False):
                key = f"out_thr{'N' if thr is None else '1'}_w{int(uw)}_s{int(sm)}"
                raw = g["raw"].copy()
                out = to_3channel(raw, items, thr, uw, sm)
                assert out.dtype == np.uint8 and np.array_equal(out, g[key]), key
                assert np.array_equal(raw, g["raw"])        # input untouched


def test_argmax_large_vs_oracle(cuda):
    from cvmhot.common.utils.image import class_ids, to_3channel
    from cvmhot.data.label_spec import SEMSEG_CLASS_MAPPING
    rng = np.random.default_rng(3)
    x = rng.normal(0, 1, (2, 64, 96, 20)).astype(np.float32)
    x[0, 0, 0, 14:19] = 1.0                                   # tie -> first
    xd = torch.from_numpy(x).to(cuda)
    ids = class_ids(xd, 14, 5).cpu().numpy()
    assert np.array_equal(ids, image_np.class_ids(x[..., 14:], 5))
    cols = list(SEMSEG_CLASS_MAPPING.values())
    bgr = to_3channel(xd[1, :, :, 14:19], SEMSEG_CLASS_MAPPING, 0.25, True, True).cpu().numpy()
    assert np.array_equal(bgr, image_np.to_3channel(x[1, :, :, 14:19], cols, 0.25, True, True))
    one = to_3channel(xd[0, :, :, 3:4], [("x", (10, 200, 30))], 0.1, True, False).cpu().numpy()      # binary case
    assert np.array_equal(one, image_np.to_3channel(x[0, :, :, 3:4], [(10, 200, 30)], 0.1, True, False))
