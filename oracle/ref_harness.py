"""TEST INFRASTRUCTURE ONLY — drives the reference's OWN functions (oracle/ref_import.py) on synthetic inputs.

Used by tests/golden/make_golden.py (golden vectors), the live-reference tests and bench.py's reference arm (the CPU
implementation of the path timed on the box's host cores).  Everything that computes here is reference code:

  render  ProcessImages.calc_img_data + the scatter lines + fill_heatmap (numba), driven by the loop of
          models/centernet/processor.py:264-334 (restated here only as far as the control flow goes: the image decode /
          augmentation part of process() needs a database document and albumentations)
  loss    CenternetLoss / CentertrackerLoss (models/centernet/loss.py:6-155, models/centertracker/loss.py:7-28) executed over
          oracle/tf_shim.py (TF's ops on torch-CPU fp32, all host threads)
  decode  process_2d_output (models/centernet/post_processing.py:6-66) - Profile R only; the reference has no top-K decode

Profile N (per-class heatmaps, what BASELINE.json measures) does not exist in the reference; it is run through the same
reference lines by (a) handing fill_heatmap a channel-sliced view whose channel 0 is the object's class plane and
(b) moving the channel positions the loss constructor pre-computed (instance attributes only).
"""
import numpy as np

_POS_ATTRS = ("class_pos", "r_offset_pos", "fullbox_pos", "l_shape_pos", "radial_dist_pos", "orientation_pos",
              "obj_dims_pos", "track_offset_pos")


def ref_loss_object(r, nb, hm=1, track=False, l_shape=False, info3d=False):
    """CenternetLoss / CentertrackerLoss of the reference (models/centernet/loss.py:6-29, centertracker/loss.py:7-14).
    hm > 1 (Profile N): the reference hard-codes ONE heatmap channel (params.py:56, loss.py:12); the per-class-heatmap
    layout is run through the very same reference lines by switching the class field off and moving the channel
    positions the constructor pre-computed (instance attributes only, the class and its methods stay untouched)."""
    P = (r["CentertrackerParams"] if track else r["CenternetParams"])(nb)
    P.REGRESSION_FIELDS["l_shape"].active = l_shape
    P.REGRESSION_FIELDS["3d_info"].active = info3d
    if hm > 1:
        P.REGRESSION_FIELDS["class"].active = False
    loss = (r["CentertrackerLoss"] if track else r["CenternetLoss"])(P)
    if hm > 1:
        loss.obj_pos = [0, hm]
        for a in _POS_ATTRS:
            if hasattr(loss, a):
                v = getattr(loss, a)
                setattr(loss, a, [x + hm - 1 for x in v] if isinstance(v, list) else v + hm - 1)
    return loss, P


def render_image_ref(r, proc, params, hm, H, W, boxes, cls, ignore):
    """One image's y_true [H,W,Ct] with the reference's own functions; control flow of processor.py:264-334.
    proc: a reference ProcessImages (for calc_img_data), params: its CenternetParams, hm: heatmap channels (1 or nb)."""
    fill = r["fill_heatmap"]
    Cp = params.mask_channels() + (hm - 1)
    heat = np.zeros((H, W, Cp), dtype=np.float32)                          # :267
    weights = np.ones((H, W), dtype=np.float32)                            # :268
    so = hm - 1
    for bbox, c in zip(boxes, cls):
        center, loc_off, _, _ = proc.calc_img_data(list(bbox), None, W, H)     # :277
        gt_center = heat[center[1]][center[0]][:]
        if params.REGRESSION_FIELDS["class"].active:
            gt_center[params.start_idx("class") + int(c)] = 1.0               # :290
        gt_center[so + params.start_idx("r_offset"):so + params.end_idx("r_offset")] = loc_off      # :292
        gt_center[so + params.start_idx("fullbox"):so + params.end_idx("fullbox")] = [bbox[2], bbox[3]]   # :294
        plane = heat[:, :, int(c):] if hm > 1 else heat                      # fill_heatmap writes channel 0 of what it is given
        fill(plane, params.VARIANCE_ALPHA, params.R, weights, center[0], center[1], bbox[2], bbox[3], W, H)   # :302
    for ia in ignore:                                                        # :318-323
        weights[int(ia[1]):int(ia[1] + ia[3]), int(ia[0]):int(ia[0] + ia[2])] = 0.0
    return np.concatenate((heat, np.expand_dims(weights, axis=-1)), axis=-1)   # :334


def make_render_ctx(r, nb, hm, H, W):
    P = r["CenternetParams"](nb)
    P.INPUT_HEIGHT, P.INPUT_WIDTH = H * P.R, W * P.R
    if hm > 1:
        P.REGRESSION_FIELDS["class"].active = False
    return r["ProcessImages"](P), P
