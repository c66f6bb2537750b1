"""CenterTracker loss = CenterNet loss + MSE on the track offset at the peaks (reference models/centertracker/loss.py:6-28).
The track term is just one more field of the fused pass (the layout carries it), so `call` needs no extra work."""
from cvmhot.models.centernet.loss import CenternetLoss


class CentertrackerLoss(CenternetLoss):
    def __init__(self, params, process_group=None):
        super().__init__(params, process_group)
        if not params.REGRESSION_FIELDS["track_offset"].active:
            raise AssertionError("CenterTracker needs the track_offset field")      # reference :13-14
        self.track_offset_pos = [params.start_idx("track_offset"), params.end_idx("track_offset")]

    def track_offset_loss(self, y_true, y_pred):
        return self._term(y_true, y_pred, "track_offset")
