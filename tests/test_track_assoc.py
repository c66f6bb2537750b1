"""CenterTrack association: oracle hand cases on CPU, cvm_track_associate against the oracle on
the GPU (bit-exact indices), and the Tracker mirror over a short synthetic sequence.  The reference has no association code,
so these cases pin the published algorithm (oracle/track_np.py header)."""
import numpy as np
import pytest
import torch

from oracle import track_np


def _img(cent, sizes, cls, scores=None, track=None):
    cent = np.asarray(cent, np.float32).reshape(-1, 2)
    K = cent.shape[0]
    sizes = np.asarray(sizes, np.float32).reshape(-1, 2)
    boxes = np.concatenate([cent - sizes / 2, sizes], axis=1).astype(np.float32)
    scores = np.linspace(0.9, 0.5, K).astype(np.float32) if scores is None else np.asarray(scores, np.float32)
    track = np.zeros((K, 2), np.float32) if track is None else np.asarray(track, np.float32)
    return cent, track, boxes, scores, np.asarray(cls, np.int32)


def test_oracle_hand_cases():
    prev_c = np.array([[10, 10], [50, 50], [12, 10]], np.float32)
    prev_s = np.array([[8, 8], [8, 8], [8, 8]], np.float32)
    prev_cls = np.array([0, 0, 1], np.int32)
    # closest-of-class wins; the class-1 detection can only take track 2; far detection stays unmatched
    c, t, b, s, k = _img([[11, 10], [11, 10], [52, 50], [200, 200]], [[8, 8]] * 4, [0, 1, 0, 0])
    assert track_np.associate_image(c, t, b, s, k, prev_c, prev_s, prev_cls).tolist() == [0, 2, 1, -1]
    # greedy order: the first (higher score) detection takes the track although the second is closer
    c, t, b, s, k = _img([[13, 10], [10, 10]], [[8, 8]] * 2, [0, 0])
    assert track_np.associate_image(c, t, b, s, k, prev_c[:1], prev_s[:1], prev_cls[:1]).tolist() == [0, -1]
    # the tracking offset moves the detection back onto the previous centre
    c, t, b, s, k = _img([[30, 10]], [[8, 8]], [0], track=[[-20, 0]])
    assert track_np.associate_image(c, t, b, s, k, prev_c, prev_s, prev_cls).tolist() == [0]
    # radius: squared distance must not exceed either box area (64 here): 8 px away is allowed, 8.1 is not
    c, t, b, s, k = _img([[18, 10], [58.1, 50]], [[8, 8]] * 2, [0, 0])
    assert track_np.associate_image(c, t, b, s, k, prev_c, prev_s, prev_cls).tolist() == [0, -1]
    # a small detection limits the radius too (area 4: only within 2 px)
    c, t, b, s, k = _img([[13, 10]], [[2, 2]], [0])
    assert track_np.associate_image(c, t, b, s, k, prev_c, prev_s, prev_cls).tolist() == [-1]
    # tie in distance: lowest previous index (numpy argmin)
    pc = np.array([[10, 10], [14, 10]], np.float32)
    c, t, b, s, k = _img([[12, 10]], [[8, 8]], [0])
    assert track_np.associate_image(c, t, b, s, k, pc, prev_s[:2], np.zeros(2, np.int32)).tolist() == [0]
    # score threshold and NaN scores skip the detection; no previous tracks at all
    c, t, b, s, k = _img([[10, 10], [50, 50]], [[8, 8]] * 2, [0, 0], scores=[0.2, np.nan])
    assert track_np.associate_image(c, t, b, s, k, prev_c, prev_s, prev_cls, min_score=0.3).tolist() == [-1, -1]
    assert track_np.associate_image(c, t, b, s, k, prev_c[:0], prev_s[:0], prev_cls[:0]).tolist() == [-1, -1]


def _scene(rng, B, K, M, n_cls=4):
    prev_c = rng.uniform(0, 400, (B, M, 2)).astype(np.float32)
    prev_s = rng.uniform(2, 60, (B, M, 2)).astype(np.float32)
    prev_cls = rng.integers(0, n_cls, (B, M)).astype(np.int32)
    cent = rng.uniform(0, 400, (B, K, 2)).astype(np.float32)
    sizes = rng.uniform(2, 60, (B, K, 2)).astype(np.float32)
    cls = rng.integers(0, n_cls, (B, K)).astype(np.int32)
    track = rng.normal(0, 3, (B, K, 2)).astype(np.float32)
    for b in range(B):            # most detections are moved previous tracks (so that matches exist), in random order
        n = min(K, M) * 3 // 4
        src = rng.permutation(M)[:n]
        dst = rng.permutation(K)[:n]
        true_prev = prev_c[b, src]
        move = rng.normal(0, 6, (n, 2)).astype(np.float32)
        cent[b, dst] = true_prev + move
        track[b, dst] = -move + rng.normal(0, 1.0, (n, 2)).astype(np.float32)
        sizes[b, dst] = prev_s[b, src] * rng.uniform(0.8, 1.2, (n, 2)).astype(np.float32)
        cls[b, dst] = prev_cls[b, src]
        if b % 3 == 0 and n > 4:  # exact duplicates: distance ties between previous tracks
            prev_c[b, src[1]] = prev_c[b, src[0]]
            prev_cls[b, src[1]] = prev_cls[b, src[0]]
    scores = np.sort(rng.uniform(0, 1, (B, K)).astype(np.float32), axis=1)[:, ::-1].copy()
    scores[0, K // 2] = np.nan
    boxes = np.concatenate([cent - sizes / 2, sizes], axis=2).astype(np.float32)
    return dict(centers=cent, track=track, boxes=boxes, scores=scores, cls=cls), prev_c, prev_s, prev_cls


@pytest.mark.gpu
@pytest.mark.parametrize("B,K,M,seed", [(8, 100, 100, 0), (5, 100, 37, 1), (3, 17, 300, 2), (2, 1, 1, 3), (4, 64, 0, 4)])
def test_associate_vs_oracle(cuda, B, K, M, seed):
    from cvmhot import ops
    rng = np.random.default_rng(seed)
    det, pc, ps, pk = _scene(rng, B, K, max(M, 1))
    pc, ps, pk = pc[:, :M], ps[:, :M], pk[:, :M]
    cnt = rng.integers(0, M + 1, B).astype(np.int32)
    cnt[0] = M
    det_d = {k: torch.from_numpy(v).to(cuda) for k, v in det.items()}
    for count, min_score in ((None, 0.0), (cnt, 0.0), (cnt, 0.4)):
        got = ops.track_associate(det_d, torch.from_numpy(np.ascontiguousarray(pc)).to(cuda), torch.from_numpy(np.ascontiguousarray(ps)).to(cuda),
                                  torch.from_numpy(np.ascontiguousarray(pk)).to(cuda), None if count is None else torch.from_numpy(count).to(cuda),
                                  min_score).cpu().numpy()
        ref = track_np.associate(det, pc, ps, pk, count, min_score)
        assert np.array_equal(got, ref)
        if M > 0 and count is None and min_score == 0.0 and K > 10:
            assert (got >= 0).sum() > 0                                   # the scenes do contain matches
        for b in range(B):                                                # a previous track is matched at most once
            m = got[b][got[b] >= 0]
            assert len(set(m.tolist())) == len(m)


def _oracle_associate(det, prev_centers, prev_sizes, prev_cls, prev_count=None, min_score=0.0):
    """stand-in for ops.track_associate on CPU tensors (the host logic of Tracker is what the CPU test is about)"""
    d = {k: v.numpy() for k, v in det.items()}
    m = track_np.associate(d, prev_centers.numpy(), prev_sizes.numpy(), prev_cls.numpy(),
                           None if prev_count is None else prev_count.numpy(), min_score)
    return torch.from_numpy(m)


def test_tracker_host_logic_on_cpu(monkeypatch):
    """The id bookkeeping of the Tracker mirror (births, compaction of the kept tracks, id continuity) with the matcher
    replaced by the oracle, so that it runs without a GPU."""
    from cvmhot import ops
    monkeypatch.setattr(ops, "track_associate", _oracle_associate)
    _run_tracker_sequence(torch.device("cpu"))


@pytest.mark.gpu
def test_tracker_keeps_ids_over_frames(cuda):
    _run_tracker_sequence(cuda)


def _run_tracker_sequence(cuda):
    """Three frames of objects drifting by a known motion: decode-like dicts with exact tracking offsets -> the ids of
    frame 0 survive, an object that appears later gets a fresh id, low-score clutter stays untracked."""
    from cvmhot.models.centertracker import Tracker
    rng = np.random.default_rng(5)
    B, K, n = 2, 16, 6
    pos = rng.uniform(60, 300, (B, n, 2)).astype(np.float32)
    pos[:, :, 0] += np.arange(n, dtype=np.float32)[None, :] * 40          # well separated
    vel = rng.uniform(-4, 4, (B, n, 2)).astype(np.float32)
    tracker = Tracker(min_score=0.1, new_thresh=0.3)
    seen = []
    for f in range(3):
        n_f = n if f < 2 else n + 1                                       # a newcomer in the last frame
        cent = np.full((B, K, 2), -1000, np.float32)
        cur = pos + vel * f
        cent[:, :n] = cur
        if f == 2:
            cent[:, n] = [350, 20]
        perm = np.stack([rng.permutation(n_f) for _ in range(B)])
        cent[:, :n_f] = np.take_along_axis(cent[:, :n_f], perm[..., None], axis=1)
        track = np.zeros((B, K, 2), np.float32)
        track[:, :n_f] = np.take_along_axis(np.concatenate([-vel, np.zeros((B, 1, 2), np.float32)], axis=1)[:, :n_f], perm[..., None], axis=1)
        sizes = np.full((B, K, 2), 20, np.float32)
        scores = np.zeros((B, K), np.float32)
        scores[:, :n_f] = 0.8
        scores[:, n_f:n_f + 2] = 0.2                                       # clutter below new_thresh
        det = dict(centers=cent, track=track, boxes=np.concatenate([cent - sizes / 2, sizes], 2).astype(np.float32), scores=scores,
                   cls=np.zeros((B, K), np.int32))
        ids = tracker.step({k: torch.from_numpy(v).to(cuda) for k, v in det.items()}).cpu().numpy()
        inv = np.argsort(perm, axis=1)                                    # ids in object order
        seen.append(np.take_along_axis(ids[:, :n_f], inv, axis=1))
        assert (ids[:, n_f:] == 0).all()
    assert (seen[0] > 0).all() and all(len(set(r.tolist())) == n for r in seen[0])
    assert np.array_equal(seen[1], seen[0])
    assert np.array_equal(seen[2][:, :n], seen[0])
    assert (seen[2][:, n] == n + 1).all()


@pytest.mark.gpu
def test_decode_to_tracker_pipeline(cuda):
    """The documented usage end to end: decode_topk of a CenterTracker layout -> Tracker.step over three frames.  Objects
    move by a known velocity, the track_offset head predicts (previous centre - centre) in input px as the processor
    scatters it (centertracker/processor.py:82-89): ids must survive, `centers + track` must land on the previous centres."""
    from cvmhot.models.centertracker import CentertrackerParams, Tracker
    from cvmhot.models.centernet.post_processing import decode_topk
    H, W, nb, B, n = 64, 96, 4, 2, 5
    p = CentertrackerParams(nb, True)
    p.INPUT_HEIGHT, p.INPUT_WIDTH = H * 2, W * 2
    from cvmhot.layout import layout_from_params
    L = layout_from_params(p)
    rng = np.random.default_rng(8)
    pos = np.stack([np.stack([np.linspace(12, W - 12, n), rng.uniform(12, H - 12, n)], 1) for _ in range(B)])   # mask px
    vel = rng.integers(-2, 3, (B, n, 2)).astype(np.float64)
    cls = rng.integers(0, nb, (B, n))
    tracker = Tracker(min_score=0.1, new_thresh=0.3)
    all_ids, prev_centres = [], None
    for f in range(3):
        yp = np.zeros((B, H, W, L.Cp), np.float32)
        yp[..., :nb] = 0.01
        cur = pos + vel * f
        for b in range(B):
            for i in range(n):
                x, y = int(cur[b, i, 0]), int(cur[b, i, 1])
                yp[b, y, x, cls[b, i]] = 0.9 - 0.05 * i
                yp[b, y, x, L.off_roff:L.off_roff + 2] = 0.0
                yp[b, y, x, L.off_box:L.off_box + 2] = 16.0
                yp[b, y, x, L.off_track:L.off_track + 2] = -vel[b, i] * 2.0      # input px (R = 2)
        det = decode_topk(torch.from_numpy(yp).to(cuda), p, K=16)
        ids = tracker.step(det).cpu().numpy()
        centres = det["centers"].cpu().numpy()[:, :n]
        if prev_centres is not None:
            q = centres + det["track"].cpu().numpy()[:, :n]
            assert np.allclose(q, prev_centres)                                   # offset convention: centre + track = previous centre
        prev_centres = centres
        assert (ids[:, :n] > 0).all() and (ids[:, n:] == 0).all()
        all_ids.append(ids[:, :n].copy())
    assert np.array_equal(all_ids[0], all_ids[1]) and np.array_equal(all_ids[1], all_ids[2])
