"""CenterTracker side of the hot path: parameters, the previous-frame heatmap / track-offset pre-processing, the loss with
its track-offset term, and the association of the decoded tracking offsets (which the reference does not have)."""
from .loss import CentertrackerLoss
from .params import CentertrackerParams
from .processor import CenterTrackerProcess
from .tracking import Tracker

__all__ = ["CentertrackerLoss", "CentertrackerParams", "CenterTrackerProcess", "Tracker"]
