"""TEST INFRASTRUCTURE ONLY — channel layout used by the CPU oracle.

Restates the channel bookkeeping of the reference (models/centernet/params.py:45-77):
channel(s) [0, hm) hold the heatmap (the reference hard-codes one objectness channel,
`start_idx = 1`, params.py:56), then the active regression fields follow in OrderedDict order
class -> r_offset -> fullbox -> l_shape -> 3d_info -> track_offset (params.py:45-52).
`y_true` carries one extra trailing channel, the loss-weights plane (processor.py:334).

Two profiles share this description:
  Profile R: hm = 1, class field active   (bit-for-bit the current reference)
  Profile N: hm = nb_classes, class field off (canonical CenterNet, what BASELINE.json measures)
"""
from dataclasses import dataclass, field
from typing import List

# loss kinds of CenternetLoss.calc_loss (models/centernet/loss.py:115-124)
KIND_MSE, KIND_MAE, KIND_MAPE, KIND_CE = 0, 1, 2, 3
# post transform applied to the reduced value (orientation_loss, loss.py:98)
POST_NONE, POST_ORIENT = 0, 1


@dataclass
class Field:
    name: str
    off: int
    size: int
    kind: int
    weight: float
    post: int = POST_NONE


@dataclass
class Layout:
    H: int
    W: int
    hm: int                     # number of leading heatmap channels
    nb_classes: int
    Cp: int                     # channels of y_pred
    Ct: int                     # channels of y_true (= Cp + 1, weights plane last)
    R: float = 2.0
    alpha: float = 0.9          # VARIANCE_ALPHA (params.py:35)
    focal_a: float = 2.0        # FOCAL_LOSS_ALPHA (params.py:40)
    focal_b: float = 4.0        # FOCAL_LOSS_BETA  (params.py:41)
    off_class: int = -1
    off_roff: int = -1
    off_box: int = -1
    off_track: int = -1
    fields: List[Field] = field(default_factory=list)   # loss terms in the order of loss.py:142-153


def make_layout(H, W, nb_classes, profile="N", track=False, l_shape=False, info3d=False,
                R=2.0, alpha=0.9, focal_a=2.0, focal_b=4.0, class_field=None) -> Layout:
    """Build the layout the way params.start_idx/end_idx would (params.py:54-77)."""
    hm = nb_classes if profile == "N" else 1
    if class_field is None:
        class_field = (profile == "R")
    idx = hm
    fields = []
    L = Layout(H=H, W=W, hm=hm, nb_classes=nb_classes, Cp=0, Ct=0, R=R, alpha=alpha,
               focal_a=focal_a, focal_b=focal_b)
    if class_field:
        L.off_class = idx
        fields.append(Field("class", idx, nb_classes, KIND_CE, 0.5))          # loss.py:62-66,143
        idx += nb_classes
    L.off_roff = idx
    fields.append(Field("r_offset", idx, 2, KIND_MAE, 0.2))                    # loss.py:68-72,145
    idx += 2
    L.off_box = idx
    fields.append(Field("fullbox", idx, 2, KIND_MAE, 0.1))                     # loss.py:74-78,147
    idx += 2
    if l_shape:
        fields.append(Field("l_shape", idx, 7, KIND_MSE, 0.1))                 # loss.py:80-84,149
        idx += 7
    if info3d:
        # split exactly like loss.py:24-29: radial (mape), orientation (mae + post), dims (mse)
        fields.append(Field("radial_dist", idx, 1, KIND_MAPE, 0.1))            # loss.py:86-90,151
        fields.append(Field("orientation", idx + 1, 1, KIND_MAE, 0.2, POST_ORIENT))  # loss.py:92-99,152
        fields.append(Field("obj_dims", idx + 2, 3, KIND_MSE, 0.1))            # loss.py:101-105,153
        idx += 5
    if track:
        L.off_track = idx
        fields.append(Field("track_offset", idx, 2, KIND_MSE, 0.1))            # centertracker/loss.py:16-27
        idx += 2
    L.Cp = idx
    L.Ct = idx + 1
    L.fields = fields
    return L
