"""Class tables used by the hot path (names, order and BGR display colours as in the reference's data/label_spec.py:6-30)."""
from collections import OrderedDict

OD_CLASS_MAPPING = OrderedDict((
    ("car", (96, 96, 192)), ("truck", (96, 192, 192)), ("van", (128, 192, 96)),
    ("motorbike", (194, 96, 64)), ("cyclist", (64, 194, 128)), ("ped", (196, 64, 196)),
))
OD_CLASS_IDX = {name: i for i, name in enumerate(OD_CLASS_MAPPING)}

SEMSEG_CLASS_MAPPING = OrderedDict((
    ("road", (32, 32, 64)), ("lane_markings", (0, 0, 255)), ("undriveable", (96, 128, 128)),
    ("movable", (102, 255, 0)), ("ego_car", (255, 0, 204)),
))
SEMSEG_CLASS_IDX = {name: i for i, name in enumerate(SEMSEG_CLASS_MAPPING)}
