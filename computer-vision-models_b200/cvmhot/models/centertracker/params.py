"""CenterTracker parameters: CenterNet + an active track_offset field (reference models/centertracker/params.py:3-24)."""
from cvmhot.models.centernet.params import CenternetParams


class CentertrackerParams(CenternetParams):
    def __init__(self, nb_classes, per_class_heatmap: bool = False):
        super().__init__(nb_classes, per_class_heatmap)
        self.LOAD_PATH_BASE = None
        self.LOAD_PATH = None
        # simulation of the t-1 detections
        self.FN_PROB = 0.0
        self.FP_PROB = 0.0
        self.POS_NOISE_WEIGHT = 0.0
        self.REGRESSION_FIELDS["track_offset"] = CenternetParams.RegressionField(
            True, 2, 0.1, "x and y offset to track at t-1 relative to input size")

    def serialize(self):
        d = super().serialize()
        d["LOAD_PATH_BASE"] = self.LOAD_PATH_BASE
        return d
