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

    # -- host-side box bookkeeping, same rules as the reference ---------------------------------------------------------
    def clip_to_img(self, bbox, min_x, min_y, max_x, max_y):
        """[x, y, w, h] clipped to the image rectangle (reference processor.py:46-56)."""
        x, y, w, h = bbox
        x1, y1 = np.clip(x + w, min_x, max_x), np.clip(y + h, min_y, max_y)
        x0, y0 = np.clip(x, min_x, max_x), np.clip(y, min_y, max_y)
        return [x0, y0, x1 - x0, y1 - y0]

    def calc_img_data(self, box2d, box3d, mask_width, mask_height):
        """Centre pixel, sub-pixel offset and - from the 8 projected corners of the 3D box - the L-shape targets of an object
        (reference processor.py:58-115): returns (center, loc_off, valid_l_shape, [bottom_left_off, bottom_center_off,
        bottom_right_off, center_height]).  The device repeats the centre math; the L-shape is host math on 8 points."""
        R = float(self.params.R)
        x, y, w, h = np.asarray(box2d, dtype=np.float64) / R
        cxf, cyf = x + float(w) / 2.0, y + float(h) / 2.0
        cx = max(0, min(mask_width - 1, int(cxf)))
        cy = max(0, min(mask_height - 1, int(cyf)))
        center, loc_off = [cx, cy], [cxf - cx, cyf - cy]
        if box3d is None:
            return center, loc_off, False, []
        pts = np.asarray(box3d, dtype=np.float64).reshape(8, 2)
        top, bottom = pts[[0, 3, 4, 7]], pts[[1, 2, 5, 6]]            # corner order of the reference (:76-77)
        i_left, i_right = int(np.argmin(bottom[:, 0])), int(np.argmax(bottom[:, 0]))   # first extremum, like np.argmin/argmax
        rest = [i for i in range(4) if i not in (i_left, i_right)]    # the masked-array argmax of :83-87: first maximum in y
        i_mid = max(rest, key=lambda i: bottom[i, 1])
        bottom_left, bottom_right, bottom_center = bottom[i_left], bottom[i_right], bottom[i_mid]
        if bottom_center[1] < min(bottom_left[1], bottom_right[1]):   # :90-93 (non-convex projection)
            center_height = box2d[3]
            bottom_center = bottom_right
        else:
            center_height = bottom_center[1] - top[i_mid][1]
        cp = np.asarray([cxf, cyf]) * R                                # :98
        offs = [bottom_left - cp, bottom_center - cp, bottom_right - cp]
        to_center = np.asarray([R * mask_width * 0.5, R * mask_height * 0.5])
        valid = all(np.sum(np.abs(o - to_center)) <= R * mask_width * 2 for o in (offs[0], offs[2], offs[1]))   # :107-115
        return center, loc_off, valid, [offs[0], offs[1], offs[2], center_height]

    def extra_targets(self, obj, bbox, mask_width, mask_height):
        """The l_shape (7) / 3d_info (5) regression targets of one kept object in channel order (reference :296-299), or
        None when the L-shape is active but not valid: the object then becomes an ignore area (:283-285)."""
        fields = self.params.REGRESSION_FIELDS
        vals = []
        if fields["l_shape"].active:
            box3d = obj["box3d"] if obj.get("box3d_valid") else None
            _, _, valid, ls = self.calc_img_data(bbox, box3d, mask_width, mask_height)
            if not valid:
                return None
            vals += [*ls[0], *ls[1], *ls[2], ls[3]]
        if fields["3d_info"].active:
            import math
            vals += [math.sqrt(obj["x"] ** 2 + obj["y"] ** 2 + obj["z"] ** 2), obj["orientation"], obj["width"], obj["height"],
                     obj["length"]]
        return vals

    def filter_objects(self, objects, img_w, img_h, extras=None):
        """clip + MIN_BOX_AREA filter (reference processor.py:241-253): kept boxes, their class ids, ignore areas.  With the
        l_shape / 3d_info fields active, `extras` (a list) receives each kept object's extra targets, and objects without a
        valid L-shape join the ignore areas (:283-285)."""
        p = self.params
        want_extra = extras is not None and (p.REGRESSION_FIELDS["l_shape"].active or p.REGRESSION_FIELDS["3d_info"].active)
        mw, mh = p.INPUT_WIDTH // p.R, p.INPUT_HEIGHT // p.R
        boxes, cls, ignore, late_ignore = [], [], [], []
        for obj in objects:
            cb = self.clip_to_img(obj["box2d"], 0, 0, img_w, img_h)
            if cb[2] * cb[3] > p.MIN_BOX_AREA:
                if want_extra:
                    ex = self.extra_targets(obj, cb, mw, mh)
                    if ex is None:
                        late_ignore.append(cb)
                        continue
                    extras.append(ex)
                boxes.append(cb)
                c = obj["obj_class"]
                cls.append(OD_CLASS_IDX[c] if isinstance(c, str) else int(c))
            else:
                ignore.append(cb)
        return boxes, cls, ignore + late_ignore

    def extra_layout(self):
        """(first channel, channel count) of the l_shape / 3d_info block in y_true, (0, 0) when neither is active."""
        p = self.params
        keys = [k for k in ("l_shape", "3d_info") if p.REGRESSION_FIELDS[k].active]
        if not keys:
            return 0, 0
        return p.start_idx(keys[0]), sum(p.REGRESSION_FIELDS[k].size for k in keys)

    # -- fast path ------------------------------------------------------------------------------------------------------
    def render_batch(self, samples, out=None, extra_ignore=None):
        """samples: list of dicts with key "objects" (list of {"box2d": [x,y,w,h], "obj_class": name or id}).
        Returns y_true [B,H,W,Ct] float32 on the device (reference processor.py:264-334 for every sample at once)."""
        p = self.params
        L = layout_from_params(p)
        b_boxes, b_cls, b_ign = [], [], []
        ex_off, ex_n = self.extra_layout()
        extras = [] if ex_n else None
        for i, s in enumerate(samples):
            boxes, cls, ign = self.filter_objects(s["objects"], p.INPUT_WIDTH, p.INPUT_HEIGHT, extras)
            if extra_ignore is not None:
                ign = ign + list(extra_ignore[i])
            b_boxes.append(boxes)
            b_cls.append(cls)
            b_ign.append(ign)
        extra = np.asarray(extras, dtype=np.float32).reshape(-1, ex_n) if ex_n else None
        return self.render_packed(L, *pack_objects(b_boxes, b_cls), *pack_boxes(b_ign), out=out, extra=extra, extra_off=ex_off)

    def render_raw_batch(self, raw_boxes, raw_cls, raw_offsets, out=None, raw_track=None):
        """Device-resident raw labels -> y_true, no host loop: raw_boxes [n,4] float64 CUDA (x, y, w, h in input px, NOT
        yet clipped), raw_cls [n] int32 CUDA (OD_CLASS_IDX), raw_offsets [B+1] int32 CUDA.  The clip + MIN_BOX_AREA filter
        of the reference (processor.py:46-56,241-253) runs on the device (cvm_prepare_objects), then the render."""
        p = self.params
        L = layout_from_params(p)
        objs, offs, ign, ioffs = ops.prepare_objects(raw_boxes, raw_cls, raw_offsets, p.INPUT_WIDTH, p.INPUT_HEIGHT,
                                                     p.MIN_BOX_AREA, raw_track)
        return ops.render_gt(L, objs, offs, int(raw_offsets.numel()) - 1, ign, ioffs, out=out)

    def render_packed(self, L, rec, offsets, ign_rec, ign_offsets, out=None, extra=None, extra_off=0):
        dev = self.device
        B = len(offsets) - 1
        objs_d = ops.to_device_records(rec, ops.OBJ_DTYPE, dev)
        offs_d = torch.from_numpy(offsets).to(dev, non_blocking=True)
        ign_d = ops.to_device_records(ign_rec, ops.BOX_DTYPE, dev)
        ioffs_d = torch.from_numpy(ign_offsets).to(dev, non_blocking=True)
        extra_d = torch.from_numpy(np.ascontiguousarray(extra)).to(dev, non_blocking=True) if extra is not None and len(extra) else None
        return ops.render_gt(L, objs_d, offs_d, B, ign_d, ioffs_d, out=out, extra=extra_d, extra_off=extra_off)

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
