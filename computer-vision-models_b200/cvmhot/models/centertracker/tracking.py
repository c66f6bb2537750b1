"""CenterTrack association on the device.

The reference trains the tracking head (models/centertracker/loss.py:16-28, processor.py:77-89) but never wrote its
consumer (models/centertracker/__init__.py:5 exports params, processor and loss only).  This is the published greedy
matcher (Zhou et al., "Tracking Objects as Points", ECCV 2020, src/lib/utils/tracker.py) on top of the top-K decode:
`Tracker.step` takes the decode_topk dict of a batch of frames (one independent video stream per batch row) and returns the
track id of every detection.  Matching runs in libcvmhot (cvm_track_associate); the id bookkeeping is a few tensor ops.
"""
import torch

from ... import ops


class Tracker:
    """One tracker per batch row.  new_thresh: detections at or above this score that found no previous track start a new
    one (lower ones get id 0 = untracked); min_score: detections below it are ignored altogether."""

    def __init__(self, min_score=0.0, new_thresh=0.3):
        self.min_score = float(min_score)
        self.new_thresh = float(new_thresh)
        self.reset()

    def reset(self):
        self.prev = None          # dict(centers [B,K,2], sizes [B,K,2], cls [B,K], ids [B,K], count [B])
        self.next_id = None       # [B] int64

    def step(self, det):
        """det: dict of ops.decode_topk for the current frames.  Returns ids [B,K] int64 (0 = untracked)."""
        scores, boxes = det["scores"], det["boxes"]
        B, K = scores.shape
        dev = scores.device
        if self.prev is None:
            self.next_id = torch.ones(B, dtype=torch.int64, device=dev)
            match = torch.full((B, K), -1, dtype=torch.int32, device=dev)
            prev_ids = torch.zeros((B, 1), dtype=torch.int64, device=dev)
        else:
            match = ops.track_associate(det, self.prev["centers"], self.prev["sizes"], self.prev["cls"], self.prev["count"],
                                        self.min_score)
            prev_ids = self.prev["ids"]
        matched = match >= 0
        ids = torch.where(matched, torch.gather(prev_ids, 1, match.clamp(min=0).long()), torch.zeros_like(match, dtype=torch.int64))
        born = (~matched) & (scores >= self.new_thresh) & (scores >= self.min_score)
        rank = torch.cumsum(born.long(), dim=1) - 1                       # order of birth inside the frame
        ids = torch.where(born, self.next_id[:, None] + rank, ids)
        self.next_id = self.next_id + born.sum(dim=1)
        # the tracks the next frame can match: every detection that carries an id, compacted to the front
        keep = ids > 0
        order = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)
        take = lambda t: torch.gather(t, 1, order if t.dim() == 2 else order[..., None].expand(-1, -1, t.shape[-1]))
        self.prev = dict(centers=take(det["centers"]).contiguous(), sizes=take(boxes[..., 2:4]).contiguous(),
                         cls=take(det["cls"]).contiguous(), ids=take(ids), count=keep.sum(dim=1).to(torch.int32))
        return ids
