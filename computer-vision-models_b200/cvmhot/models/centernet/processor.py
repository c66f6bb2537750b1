"""Ground-truth generation for CenterNet on the GPU.

Drop-in surface of the reference's models/centernet/processor.py: `fill_heatmap` (:17-38) and
`ProcessImages(params, start_augmentation, show_debug_img).process(...)` (:41, :215-339) keep their signatures; the
tensor math (gaussian splat, weights, centre scatter, ignore areas) runs in cvm_render_gt / cvm_fill_heatmap_inplace.
The fast path is `ProcessImages.render_batch`, which renders a whole batch straight into a device tensor so that
y_true never exists on the host.
"""
import numpy as np
import torch

from cvmhot import _lib, ops
from cvmhot.layout import layout_from_params
from cvmhot.common.processors import IPreProcessor
from cvmhot.data.label_spec import OD_CLASS_IDX


def fill_heatmap(ground_truth, alpha, R, weights, center_x, center_y, width, height, mask_width, mask_height, peak=1.0):
    """Same contract as the reference (processor.py:18-22): max-combines one gaussian blob into channel 0 of
    `ground_truth` [H,W,C] and min-combines the loss weights into `weights` [H,W], IN PLACE, returns None.
    numpy arrays make a host round trip (use render_batch for throughput); CUDA tensors are updated on the device."""
    rec = np.zeros(1, dtype=ops.OBJ_DTYPE)
    rec["w"], rec["h"] = float(width), float(height)
    rec["cx"], rec["cy"] = int(center_x), int(center_y)
    rec["peak"] = float(peak)
    rec["flags"] = _lib.OBJ_EXPLICIT_CENTER | _lib.OBJ_NO_SCATTER
    if isinstance(ground_truth, torch.Tensor):
        dev = ground_truth.device
        ops.fill_heatmap_inplace(ops.to_device_records(rec, ops.OBJ_DTYPE, dev), 1, ground_truth, weights,
                                 mask_height, mask_width, R, alpha)
        return None
    dev = torch.device("cuda")
    heat = torch.from_numpy(np.ascontiguousarray(ground_truth[:, :, 0], dtype=np.float32)).to(dev)
    wts = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float32)).to(dev) if weights is not None else None
    ops.fill_heatmap_inplace(ops.to_device_records(rec, ops.OBJ_DTYPE, dev), 1, heat, wts, mask_height, mask_width, R, alpha)
    ground_truth[:, :, 0] = heat.cpu().numpy()
    if weights is not None:
        weights[...] = wts.cpu().numpy()
    return None


def pack_objects(per_image_boxes, per_image_cls, per_image_track=None):
    """Lists (one entry per image) of [n,4] fp64 boxes (x,y,w,h input px) and [n] class ids -> (records, offsets)."""
    counts = [len(b) for b in per_image_boxes]
    offsets = np.zeros(len(counts) + 1, dtype=np.int32)
    np.cumsum(counts, out=offsets[1:])
    rec = np.zeros(int(offsets[-1]), dtype=ops.OBJ_DTYPE)
    if offsets[-1] > 0:
        boxes = np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1, 4) for b in per_image_boxes if len(b)], axis=0)
        rec["x"], rec["y"], rec["w"], rec["h"] = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
        rec["cls"] = np.concatenate([np.asarray(c, dtype=np.int32).reshape(-1) for c in per_image_cls if len(c)])
        rec["peak"] = 1.0
        if per_image_track is not None:
            rec["track"] = np.concatenate([np.asarray(t, dtype=np.float32).reshape(-1, 2) for t in per_image_track if len(t)])
    return rec, offsets


def pack_boxes(per_image_boxes):
    counts = [len(b) for b in per_image_boxes]
    offsets = np.zeros(len(counts) + 1, dtype=np.int32)
    np.cumsum(counts, out=offsets[1:])
    rec = np.zeros(int(offsets[-1]), dtype=ops.BOX_DTYPE)
    if offsets[-1] > 0:
        boxes = np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1, 4) for b in per_image_boxes if len(b)], axis=0)
        rec["x"], rec["y"], rec["w"], rec["h"] = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    return rec, offsets


class ProcessImages(IPreProcessor):
    def __init__(self, params, start_augmentation=None, show_debug_img: bool = False, device=None):
        self.params = params
        self.start_augmentation = start_augmentation
        self.show_debug_img = show_debug_img
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        if params.REGRESSION_FIELDS["l_shape"].active or params.REGRESSION_FIELDS["3d_info"].active:
            raise NotImplementedError("render of l_shape / 3d_info targets is not part of the 2D heatmap path")

    # -- host-side box bookkeeping, same rules as the reference ---------------------------------------------------------
    def clip_to_img(self, bbox, min_x, min_y, max_x, max_y):
        """[x, y, w, h] clipped to the image rectangle (reference processor.py:46-56)."""
        x, y, w, h = bbox
        x1, y1 = np.clip(x + w, min_x, max_x), np.clip(y + h, min_y, max_y)
        x0, y0 = np.clip(x, min_x, max_x), np.clip(y, min_y, max_y)
        return [x0, y0, x1 - x0, y1 - y0]

    def calc_img_data(self, box2d, box3d, mask_width, mask_height):
        """Centre pixel and sub-pixel offset of a box (reference processor.py:58-67); the device does the same math."""
        if box3d is not None:
            raise NotImplementedError("l_shape targets are outside the 2D heatmap path")
        x, y, w, h = np.asarray(box2d, dtype=np.float64) / float(self.params.R)
        cxf, cyf = x + float(w) / 2.0, y + float(h) / 2.0
        cx = max(0, min(mask_width - 1, int(cxf)))
        cy = max(0, min(mask_height - 1, int(cyf)))
        return [cx, cy], [cxf - cx, cyf - cy], False, []

    def filter_objects(self, objects, img_w, img_h):
        """clip + MIN_BOX_AREA filter (reference processor.py:241-253): kept boxes, their class ids, ignore areas."""
        boxes, cls, ignore = [], [], []
        for obj in objects:
            cb = self.clip_to_img(obj["box2d"], 0, 0, img_w, img_h)
            if cb[2] * cb[3] > self.params.MIN_BOX_AREA:
                boxes.append(cb)
                c = obj["obj_class"]
                cls.append(OD_CLASS_IDX[c] if isinstance(c, str) else int(c))
            else:
                ignore.append(cb)
        return boxes, cls, ignore

    # -- fast path ------------------------------------------------------------------------------------------------------
    def render_batch(self, samples, out=None, extra_ignore=None):
        """samples: list of dicts with key "objects" (list of {"box2d": [x,y,w,h], "obj_class": name or id}).
        Returns y_true [B,H,W,Ct] float32 on the device (reference processor.py:264-334 for every sample at once)."""
        p = self.params
        L = layout_from_params(p)
        b_boxes, b_cls, b_ign = [], [], []
        for i, s in enumerate(samples):
            boxes, cls, ign = self.filter_objects(s["objects"], p.INPUT_WIDTH, p.INPUT_HEIGHT)
            if extra_ignore is not None:
                ign = ign + list(extra_ignore[i])
            b_boxes.append(boxes)
            b_cls.append(cls)
            b_ign.append(ign)
        return self.render_packed(L, *pack_objects(b_boxes, b_cls), *pack_boxes(b_ign), out=out)

    def render_raw_batch(self, raw_boxes, raw_cls, raw_offsets, out=None, raw_track=None):
        """Device-resident raw labels -> y_true, no host loop: raw_boxes [n,4] float64 CUDA (x, y, w, h in input px, NOT
        yet clipped), raw_cls [n] int32 CUDA (OD_CLASS_IDX), raw_offsets [B+1] int32 CUDA.  The clip + MIN_BOX_AREA filter
        of the reference (processor.py:46-56,241-253) runs on the device (cvm_prepare_objects), then the render."""
        p = self.params
        L = layout_from_params(p)
        objs, offs, ign, ioffs = ops.prepare_objects(raw_boxes, raw_cls, raw_offsets, p.INPUT_WIDTH, p.INPUT_HEIGHT,
                                                     p.MIN_BOX_AREA, raw_track)
        return ops.render_gt(L, objs, offs, int(raw_offsets.numel()) - 1, ign, ioffs, out=out)

    def render_packed(self, L, rec, offsets, ign_rec, ign_offsets, out=None):
        dev = self.device
        B = len(offsets) - 1
        objs_d = ops.to_device_records(rec, ops.OBJ_DTYPE, dev)
        offs_d = torch.from_numpy(offsets).to(dev, non_blocking=True)
        ign_d = ops.to_device_records(ign_rec, ops.BOX_DTYPE, dev)
        ioffs_d = torch.from_numpy(ign_offsets).to(dev, non_blocking=True)
        return ops.render_gt(L, objs_d, offs_d, B, ign_d, ioffs_d, out=out)

    # -- reference plug-in entry point -----------------------------------------------------------------------------------
    def process(self, raw_data, input_data, ground_truth, piped_params=None):
        """One sample (reference processor.py:215-339).  Returns input_data = [img float32], ground_truth [H,W,Ct] float32
        as numpy arrays, like the reference; the render itself happens on the GPU."""
        p = self.params
        if p.REGRESSION_FIELDS["track_offset"].active:
            raise AssertionError("use CenterTrackerProcess when track_offset is regressed")   # reference :231-233
        if isinstance(raw_data, list):
            raise AssertionError("a single frame is expected when no track offset is regressed")
        img = raw_data["img"]
        if isinstance(img, (bytes, bytearray, memoryview)):
            import cv2
            img = cv2.imdecode(np.frombuffer(img, np.uint8), cv2.IMREAD_COLOR)
        img = np.asarray(img)
        if img.shape[0] != p.INPUT_HEIGHT or img.shape[1] != p.INPUT_WIDTH:
            raise AssertionError("images are expected to be INPUT_HEIGHT x INPUT_WIDTH already")
        if self.start_augmentation is not None and piped_params is not None and \
                piped_params.get("epoch", 0) >= min(self.start_augmentation):
            raise NotImplementedError("image/affine augmentation (albumentations) is outside the heatmap hot path")
        y_true = self.render_batch([raw_data])
        input_data = [img.astype(np.float32)]
        ground_truth = y_true[0].cpu().numpy()
        return raw_data, input_data, ground_truth, piped_params
