"""TEST INFRASTRUCTURE ONLY - CPU restatement of the CenterTrack association (never on the product path).

PARITY UNPINNED: the reference has no association code (/root/reference/models/centertracker/__init__.py:5 exports
params / loss / processor only).  This restates the algorithm published with CenterTrack
(Zhou, Koltun, Kraehenbuehl, "Tracking Objects as Points", ECCV 2020; src/lib/utils/tracker.py, Tracker.step and
greedy_assignment) on the outputs of the Profile-N decode:
  dets = centre + tracking offset;  dist[i, j] = sum((track_j - det_i) ** 2) in fp32
  invalid = dist > track_size_j  or  dist > item_size_i  or  class_i != class_j      (sizes are box areas w*h)
  for i in score order: j = argmin(dist[i]) (first minimum); if valid: match (i, j) and remove column j
The input targets come from CenterTrackerProcess (/root/reference/models/centertracker/processor.py:77-89:
track_offset = previous centre - centre * R).
"""
import numpy as np


def associate_image(centers, track, boxes, scores, cls, prev_centers, prev_sizes, prev_cls, min_score=0.0):
    """One image.  centers/track [K,2], boxes [K,4] (tlx,tly,w,h), scores/cls [K]; prev_centers/prev_sizes [M,2], prev_cls [M].
    Returns match [K] int32 (-1 = none)."""
    f = np.float32
    centers, track, boxes = np.asarray(centers, f), np.asarray(track, f), np.asarray(boxes, f)
    prev_centers, prev_sizes = np.asarray(prev_centers, f).reshape(-1, 2), np.asarray(prev_sizes, f).reshape(-1, 2)
    K, M = centers.shape[0], prev_centers.shape[0]
    match = np.full(K, -1, np.int32)
    if M == 0:
        return match
    det = centers + track                                                     # fp32
    diff = prev_centers[None, :, :] - det[:, None, :]
    dist = diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]          # [K, M] fp32, x term + y term
    track_size = prev_sizes[:, 0] * prev_sizes[:, 1]
    item_size = boxes[:, 2] * boxes[:, 3]
    with np.errstate(invalid="ignore"):
        invalid = (dist > track_size[None, :]) | (dist > item_size[:, None]) | (np.asarray(cls)[:, None] != np.asarray(prev_cls)[None, :])
        invalid |= np.isnan(dist)
    taken = np.zeros(M, bool)
    for i in range(K):
        if not (scores[i] >= min_score):
            continue
        row = np.where(invalid[i] | taken, np.float32(np.inf), dist[i])
        j = int(np.argmin(row))
        if np.isfinite(row[j]):
            match[i] = j
            taken[j] = True
    return match


def associate(det, prev_centers, prev_sizes, prev_cls, prev_count=None, min_score=0.0):
    """Batch version over the dict of decode_np / ops.decode_topk (numpy arrays)."""
    B = det["scores"].shape[0]
    out = []
    for b in range(B):
        m = prev_centers.shape[1] if prev_count is None else int(prev_count[b])
        out.append(associate_image(det["centers"][b], det["track"][b], det["boxes"][b], det["scores"][b], det["cls"][b],
                                   prev_centers[b, :m], prev_sizes[b, :m], prev_cls[b, :m], min_score))
    return np.stack(out)
