"""The CenterNet slice of the multitask loss (reference models/multitask/loss.py:21-24,44-47).  The semseg and depth
terms are different losses outside the heatmap path.  No copy is made: the fused kernel reads the CenterNet channels in
place using the wide tensors' pixel strides."""
from cvmhot.models.centernet.loss import CenternetLoss
from cvmhot.common.utils.image import class_ids


class MultitaskLoss:
    def __init__(self, params, process_group=None):
        self.params = params
        self.cn_loss = CenternetLoss(params.cn_params, process_group)
        self.cn_offset = params.cn_offset()
        self.semseg_offset = params.semseg_offset()
        self.depth_offset = params.depth_offset()

    def calc_centernet(self, y_true, y_pred):
        a, b = self.cn_offset["y_true"]
        c, d = self.cn_offset["y_pred"]
        return self.cn_loss.call(y_true[..., a:b], y_pred[..., c:d])

    def semseg_class_ids(self, y_pred):
        """per-pixel argmax of the semseg slice (what to_3channel computes for display, multitask/callbacks.py:107-110)."""
        a, b = self.semseg_offset["y_pred"]
        return class_ids(y_pred, a, b - a)
