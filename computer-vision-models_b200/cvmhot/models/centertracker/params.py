"""CenterTracker parameters = CenterNet parameters with the track_offset head switched on, plus the knobs that simulate
the previous frame's detections (attribute surface of the reference's models/centertracker/params.py:3-24)."""
from cvmhot.models.centernet.params import CenternetParams

_TRACKER_DEFAULTS = dict(LOAD_PATH_BASE=None, LOAD_PATH=None, FN_PROB=0.0, FP_PROB=0.0, POS_NOISE_WEIGHT=0.0)


class CentertrackerParams(CenternetParams):
    def __init__(self, nb_classes, per_class_heatmap: bool = False):
        super().__init__(nb_classes, per_class_heatmap)
        for name, value in _TRACKER_DEFAULTS.items():
            setattr(self, name, value)
        self.REGRESSION_FIELDS["track_offset"].active = True

    def serialize(self):
        return dict(super().serialize(), LOAD_PATH_BASE=self.LOAD_PATH_BASE)
