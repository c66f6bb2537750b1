"""CenterNet hyper-parameters and output-channel bookkeeping (host side of `cvm_layout`).

Drop-in for the attribute surface of the reference's models/centernet/params.py:10-99: the same attribute and field
names, and start_idx / end_idx / mask_channels / serialize / save_to_storage with the same meaning, so that scripts
written against the reference keep working.  One extension: HM_CHANNELS, the number of leading heatmap channels
(1 = the reference as shipped: one objectness channel, class logits as a regression field; nb_classes with the class
field switched off = canonical CenterNet with per-class heatmaps, which is what BASELINE.json measures).
"""
import json
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any

# attribute -> default (values of params.py:20-41)
_DEFAULTS = dict(BATCH_SIZE=8, PLANED_EPOCHS=90, LOAD_WEIGHTS=None, INPUT_WIDTH=640, INPUT_HEIGHT=256, CHANNELS=3,
                 OFFSET_BOTTOM=0, MIN_BOX_AREA=15.0, R=2, VARIANCE_ALPHA=0.9, FOCAL_LOSS_ALPHA=2.0, FOCAL_LOSS_BETA=4.0,
                 CLASS_WEIGHT=1.0)

# regression heads in channel order (params.py:45-52): key, channels, loss weight(s), active by default, what it holds
_FIELDS = (
    ("class", None, 0.5, None, "class logits (one per object class)"),
    ("r_offset", 2, 0.2, True, "sub-pixel centre offset x, y"),
    ("fullbox", 2, 0.1, True, "box width, height in input pixels"),
    ("l_shape", 7, 0.1, False, "bottom-left / bottom-centre / bottom-right offsets and centre height"),
    ("3d_info", 5, [0.1, 0.2, 0.1], False, "radial distance [m], orientation [rad], width, height, length [m]"),
    ("track_offset", 2, 0.1, False, "offset x, y to the centre at t-1 in input pixels"),
)


class CenternetParams:
    @dataclass
    class RegressionField:
        active: bool = False
        size: int = 0
        loss_weight: Any = 0.0
        comment: str = ""

    def __init__(self, nb_classes: int, per_class_heatmap: bool = False):
        for name, value in _DEFAULTS.items():
            setattr(self, name, value)
        self.MASK_HEIGHT, self.MASK_WIDTH = self.INPUT_HEIGHT // self.R, self.INPUT_WIDTH // self.R
        self.NB_CLASSES = nb_classes
        self.HM_CHANNELS = nb_classes if per_class_heatmap else 1
        self.REGRESSION_FIELDS = OrderedDict()
        for key, size, weight, active, what in _FIELDS:
            if key == "class":   # the class field exists only next to a single objectness channel
                size, active = nb_classes, not per_class_heatmap
            self.REGRESSION_FIELDS[key] = CenternetParams.RegressionField(active, size, weight, what)

    # ---- channel bookkeeping: heatmap channels first, then the active fields in table order ----
    def _offsets(self):
        at, out = self.HM_CHANNELS, {}
        for key, field in self.REGRESSION_FIELDS.items():
            if field.active:
                out[key] = at
                at += field.size
        return out, at

    def start_idx(self, regression_key: str) -> int:
        offsets, _ = self._offsets()
        if regression_key not in offsets:
            raise KeyError(f"regression field {regression_key!r} is not active or does not exist")
        return offsets[regression_key]

    def end_idx(self, regression_key: str) -> int:
        return self.start_idx(regression_key) + self.REGRESSION_FIELDS[regression_key].size

    def mask_channels(self) -> int:
        return self._offsets()[1]

    def serialize(self):
        offsets, total = self._offsets()
        heads = [{"object_likelihood": {"start_idx": 0, "end_idx": self.HM_CHANNELS, "comment": "Object classes"}}]
        heads += [{key: {"start_idx": at, "end_idx": at + self.REGRESSION_FIELDS[key].size,
                         "comment": self.REGRESSION_FIELDS[key].comment}} for key, at in offsets.items()]
        return {"input": [self.INPUT_HEIGHT, self.INPUT_WIDTH, 3],
                "mask": [self.INPUT_HEIGHT // self.R, self.INPUT_WIDTH // self.R, total],
                "batch_size": self.BATCH_SIZE, "load_weights": self.LOAD_WEIGHTS, "output_fields": heads}

    def save_to_storage(self, storage_path: str):
        with open(storage_path + "/parameters.json", "w") as fh:
            json.dump(self.serialize(), fh)
