"""CenterNet hyper-parameters and output-channel bookkeeping.

Same public surface as the reference's models/centernet/params.py:10-99 (attribute names, REGRESSION_FIELDS,
start_idx / end_idx / mask_channels / serialize / save_to_storage), plus one extension: HM_CHANNELS, the number of leading
heatmap channels.  HM_CHANNELS = 1 is the reference as shipped (one objectness channel, class logits as a regression
field); HM_CHANNELS = nb_classes with the class field switched off is canonical CenterNet (per-class heatmaps).
"""
import json
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any


class CenternetParams:
    @dataclass
    class RegressionField:
        active: bool = False
        size: int = 0
        loss_weight: Any = 0.0
        comment: str = ""

    def __init__(self, nb_classes: int, per_class_heatmap: bool = False):
        F = CenternetParams.RegressionField
        # training
        self.BATCH_SIZE = 8
        self.PLANED_EPOCHS = 90
        self.LOAD_WEIGHTS = None
        # input
        self.INPUT_WIDTH = 640
        self.INPUT_HEIGHT = 256
        self.CHANNELS = 3
        self.OFFSET_BOTTOM = 0
        # box filter
        self.MIN_BOX_AREA = 15.0
        # output mask
        self.R = 2
        self.VARIANCE_ALPHA = 0.9
        self.MASK_HEIGHT = self.INPUT_HEIGHT // self.R
        self.MASK_WIDTH = self.INPUT_WIDTH // self.R
        # loss
        self.FOCAL_LOSS_ALPHA = 2.0
        self.FOCAL_LOSS_BETA = 4.0
        self.CLASS_WEIGHT = 1.0
        self.NB_CLASSES = nb_classes
        self.HM_CHANNELS = nb_classes if per_class_heatmap else 1
        self.REGRESSION_FIELDS = OrderedDict((
            ("class", F(not per_class_heatmap, nb_classes, 0.5, "Class regression")),
            ("r_offset", F(True, 2, 0.2, "x, y")),
            ("fullbox", F(True, 2, 0.1, "width, height (in [px] relative to input)")),
            ("l_shape", F(False, 7, 0.1, "bottom_left_offset, bottom_center_offset, bottom_right_offset, center_height")),
            ("3d_info", F(False, 5, [0.1, 0.2, 0.1], "radial_dist [m], orientation [rad], width, height, length [m]")),
            ("track_offset", F(False, 2, 0.1, "x and y offset to track at t-1 relative to input size")),
        ))

    def start_idx(self, regression_key: str) -> int:
        idx = self.HM_CHANNELS
        for key, f in self.REGRESSION_FIELDS.items():
            if not f.active:
                continue
            if key == regression_key:
                return idx
            idx += f.size
        raise KeyError(f"regression field {regression_key!r} is not active or does not exist")

    def end_idx(self, regression_key: str) -> int:
        return self.start_idx(regression_key) + self.REGRESSION_FIELDS[regression_key].size

    def mask_channels(self) -> int:
        return self.HM_CHANNELS + sum(f.size for f in self.REGRESSION_FIELDS.values() if f.active)

    def serialize(self):
        fields = [{"object_likelihood": {"start_idx": 0, "end_idx": self.HM_CHANNELS, "comment": "Object classes"}}]
        for key, f in self.REGRESSION_FIELDS.items():
            if f.active:
                fields.append({key: {"start_idx": self.start_idx(key), "end_idx": self.end_idx(key), "comment": f.comment}})
        return {
            "input": [self.INPUT_HEIGHT, self.INPUT_WIDTH, 3],
            "mask": [self.INPUT_HEIGHT // self.R, self.INPUT_WIDTH // self.R, self.mask_channels()],
            "batch_size": self.BATCH_SIZE,
            "load_weights": self.LOAD_WEIGHTS,
            "output_fields": fields,
        }

    def save_to_storage(self, storage_path: str):
        with open(storage_path + "/parameters.json", "w") as fh:
            json.dump(self.serialize(), fh)
