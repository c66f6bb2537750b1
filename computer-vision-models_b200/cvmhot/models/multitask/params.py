"""Multitask head = CenterNet + semseg + depth concatenated on the channel axis (reference models/multitask/params.py:7-32)."""
from cvmhot.data.label_spec import SEMSEG_CLASS_MAPPING
from cvmhot.models.centernet.params import CenternetParams


class MultitaskParams:
    def __init__(self, nb_classes, per_class_heatmap: bool = False):
        self.INPUT_WIDTH = 640
        self.INPUT_HEIGHT = 256
        self.MASK_WIDTH = self.INPUT_WIDTH // 2
        self.MASK_HEIGHT = self.INPUT_HEIGHT // 2
        self.PLANED_EPOCHS = 100
        self.BATCH_SIZE = 4
        self.cn_params = CenternetParams(nb_classes, per_class_heatmap)
        self.cn_params.CHANNELS = 3 + len(SEMSEG_CLASS_MAPPING) + 1
        self.NB_SEMSEG_CLASSES = len(SEMSEG_CLASS_MAPPING)

    # channel ranges of the concatenated tensors (reference models/multitask/loss.py:21-32)
    def cn_offset(self):
        c = self.cn_params.mask_channels()
        return {"y_true": [0, c + 1], "y_pred": [0, c]}

    def semseg_offset(self):
        cn = self.cn_offset()
        n = self.NB_SEMSEG_CLASSES
        return {"y_true": [cn["y_true"][1], cn["y_true"][1] + n + 1], "y_pred": [cn["y_pred"][1], cn["y_pred"][1] + n]}

    def depth_offset(self):
        ss = self.semseg_offset()
        return {"y_true": [ss["y_true"][1], ss["y_true"][1] + 1], "y_pred": [ss["y_pred"][1], ss["y_pred"][1] + 1]}
