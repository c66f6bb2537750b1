"""CenterTracker pre-processing: previous-frame heatmap + track-offset targets.

Follows the INTENDED behaviour of the reference's models/centertracker/processor.py:22-93 (the shipped file calls a
method that does not exist): the host keeps the random draws (false negatives / false positives /
position jitter, in the reference's draw order); the blobs are rendered by cvm_render_prev_hm.
"""
import random

import numpy as np
import torch

from cvmhot import _lib, ops
from cvmhot.layout import layout_from_params
from cvmhot.models.centernet.processor import ProcessImages, pack_objects, pack_boxes


class CenterTrackerProcess(ProcessImages):
    def random_offset(self):
        """FP displacement: +-U(10,45) mask px (reference :16-21)."""
        v = random.random() * 35 + 10
        return -v if random.random() < 0.5 else v

    def prev_heatmap_records(self, gt_2d_info):
        """gt_2d_info: iterable of [center_x, center_y, width, height].  Returns explicit-centre records (reference :26-39)."""
        recs = []
        for cx, cy, w, h in gt_2d_info:
            if random.random() > self.params.FN_PROB:
                if random.random() < self.params.FP_PROB:
                    fx, fy = int(cx + self.random_offset()), int(cy + self.random_offset())
                    recs.append((fx, fy, w, h, ((random.random() * 0.8) + 0.2) ** 2))
                nx = int(cx + np.random.normal() * self.params.POS_NOISE_WEIGHT * w)
                ny = int(cy + np.random.normal() * self.params.POS_NOISE_WEIGHT * h)
                recs.append((nx, ny, w, h, random.random() * 0.5 + 0.2))
        return recs

    @staticmethod
    def pack_prev_records(per_image_recs):
        counts = [len(r) for r in per_image_recs]
        offsets = np.zeros(len(counts) + 1, dtype=np.int32)
        np.cumsum(counts, out=offsets[1:])
        rec = np.zeros(int(offsets[-1]), dtype=ops.OBJ_DTYPE)
        i = 0
        for recs in per_image_recs:
            for cx, cy, w, h, peak in recs:
                rec[i]["cx"], rec[i]["cy"], rec[i]["w"], rec[i]["h"], rec[i]["peak"] = int(cx), int(cy), w, h, peak
                rec[i]["flags"] = _lib.OBJ_EXPLICIT_CENTER | _lib.OBJ_NO_SCATTER
                i += 1
        return rec, offsets

    def gen_prev_heatmap(self, shape, gt_2d_info, roi=None):
        """[H,W,1] previous-frame heatmap for ONE sample as a numpy array (reference :22-41)."""
        L = layout_from_params(self.params, H=int(shape[0]), W=int(shape[1]))
        rec, offs = self.pack_prev_records([self.prev_heatmap_records(gt_2d_info)])
        out = ops.render_prev_heatmap(L, ops.to_device_records(rec, ops.OBJ_DTYPE, self.device),
                                      torch.from_numpy(offs).to(self.device), 1)
        return out[0].cpu().numpy()

    def render_prev_batch(self, per_image_recs, out=None):
        """Batched fast path: list (per image) of (cx, cy, w, h, peak) -> [B,H,W,1] device tensor."""
        L = layout_from_params(self.params)
        rec, offs = self.pack_prev_records(per_image_recs)
        return ops.render_prev_heatmap(L, ops.to_device_records(rec, ops.OBJ_DTYPE, self.device),
                                       torch.from_numpy(offs).to(self.device), len(per_image_recs), out=out)

    def render_batch_with_tracks(self, samples, track_offsets, out=None):
        """Ground truth incl. the track_offset scatter (reference :77-89): track_offsets[i] is [n_kept,2] for sample i,
        already computed on the host from the t-1 transform (0,0 when the previous centre left the frame)."""
        p = self.params
        L = layout_from_params(p)
        b_boxes, b_cls, b_ign = [], [], []
        for s in samples:
            boxes, cls, ign = self.filter_objects(s["objects"], p.INPUT_WIDTH, p.INPUT_HEIGHT)
            b_boxes.append(boxes)
            b_cls.append(cls)
            b_ign.append(ign)
        rec, offs = pack_objects(b_boxes, b_cls, track_offsets)
        return self.render_packed(L, rec, offs, *pack_boxes(b_ign), out=out)

    def process(self, raw_data, input_data, ground_truth, piped_params=None):
        raise NotImplementedError(
            "the reference's CenterTrackerProcess.process is not runnable (SURVEY.md App. C.5); use "
            "render_batch_with_tracks + render_prev_batch, which implement its intended tensor math")
