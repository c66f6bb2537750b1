"""CPU: the Profile-N decode oracle (north_star's canonical CenterNet decode; SURVEY.md App. A.3.2 - no such code in the
reference, so the oracle is what the GPU path is compared with) against an independent pixel-by-pixel restatement:
3x3 peak test with in-bounds neighbours only, plateaus kept, score > 0, order (score desc, NHWC flat index asc), zero-score
tail in flat order - the tf.nn.top_k convention.  Small maps with planted ties, plateaus and border peaks."""
import numpy as np
import pytest

from oracle import decode_np
from oracle.layout import make_layout


def _naive_topk(hmap, K):
    H, W, C = hmap.shape
    cand = []
    for y in range(H):
        for x in range(W):
            for c in range(C):
                v = hmap[y, x, c]
                peak = v > 0
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        yy, xx = y + dy, x + dx
                        if (dy or dx) and 0 <= yy < H and 0 <= xx < W and hmap[yy, xx, c] > v:
                            peak = False
                cand.append((-float(v) if peak else 0.0, (y * W + x) * C + c))
    cand.sort()                                    # (-score asc = score desc, flat asc); non-peaks carry score 0
    return [(-s if s != 0.0 else 0.0, f) for s, f in cand[:K]]


@pytest.mark.parametrize("H,W,hm,K,seed", [(6, 7, 3, 20, 0), (5, 5, 1, 30, 1), (8, 4, 10, 100, 2), (3, 9, 2, 60, 3)])
def test_profile_n_decode_matches_naive(H, W, hm, K, seed):
    rng = np.random.default_rng(seed)
    L = make_layout(H, W, hm, "N")
    yp = np.zeros((H, W, L.Cp), np.float32)
    hm_vals = rng.choice(np.linspace(0.05, 0.95, 12).astype(np.float32), size=(H, W, hm))   # few distinct values: many ties
    hm_vals[rng.random((H, W, hm)) < 0.3] = 0.0                                             # zeros never count as peaks
    hm_vals[0, 0, 0] = hm_vals[H - 1, W - 1, hm - 1] = 0.99                                  # corner peaks
    hm_vals[H // 2, W // 2:W // 2 + 2, 0] = 0.97                                             # a two-pixel plateau: both kept
    yp[..., :hm] = hm_vals
    yp[..., hm:] = rng.uniform(0, 5, (H, W, L.Cp - hm)).astype(np.float32)
    got = decode_np.decode_topk_image(L, yp, K)
    want = _naive_topk(yp[..., :hm], min(K, H * W * hm))
    assert got["flat"].tolist() == [f for _, f in want]
    assert np.array_equal(got["scores"], np.array([s for s, _ in want], np.float32))
    assert np.array_equal(got["cls"], got["flat"] % hm)
    # the plateau and the corners are in
    flats = set(got["flat"].tolist())
    assert {0, ((H - 1) * W + W - 1) * hm + hm - 1} <= flats or K < 3
