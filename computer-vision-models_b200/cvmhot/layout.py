"""Channel layout of the CenterNet tensors, flattened from a params object to the POD struct of the C ABI.

Mirrors the bookkeeping of the reference's CenternetParams.start_idx/end_idx/mask_channels
(models/centernet/params.py:54-77) generalised to `hm` leading heatmap channels (the reference hard-codes 1).
"""
from dataclasses import dataclass, field
from typing import List, Tuple

from . import _lib

# (params key, loss kind, post) in the order CenternetLoss.call adds the terms (models/centernet/loss.py:142-153)
_FIELD_KINDS = {
    "class": _lib.KIND_CE, "r_offset": _lib.KIND_MAE, "fullbox": _lib.KIND_MAE, "l_shape": _lib.KIND_MSE,
    "track_offset": _lib.KIND_MSE,
}


@dataclass
class Layout:
    H: int
    W: int
    hm: int
    nb_classes: int
    Cp: int
    Ct: int
    off_class: int = -1
    off_roff: int = -1
    off_box: int = -1
    off_track: int = -1
    fields: List[Tuple[str, int, int, int, int, float]] = field(default_factory=list)  # name, off, size, kind, post, weight
    focal_a: float = 2.0
    focal_b: float = 4.0
    R: float = 2.0
    alpha: float = 0.9

    def c_struct(self) -> _lib.CvmLayout:
        s = _lib.CvmLayout()
        s.H, s.W, s.hm, s.nb_classes, s.Cp, s.Ct = self.H, self.W, self.hm, self.nb_classes, self.Cp, self.Ct
        s.off_class, s.off_roff, s.off_box, s.off_track = self.off_class, self.off_roff, self.off_box, self.off_track
        if len(self.fields) > _lib.CVM_MAX_FIELDS:
            raise ValueError("too many loss fields")
        s.n_fields = len(self.fields)
        for i, (_, off, size, kind, post, weight) in enumerate(self.fields):
            s.field_off[i], s.field_size[i], s.field_kind[i], s.field_post[i] = off, size, kind, post
            s.field_weight[i] = weight
        s.focal_a, s.focal_b, s.R, s.alpha = self.focal_a, self.focal_b, float(self.R), float(self.alpha)
        return s

    def field_names(self):
        return [f[0] for f in self.fields]


def layout_from_params(params, H=None, W=None) -> Layout:
    """Build the layout from a CenternetParams / CentertrackerParams duck-typed object.

    Extra attribute honoured: params.HM_CHANNELS (default 1 = the reference as shipped; nb_classes = canonical CenterNet).
    """
    hm = int(getattr(params, "HM_CHANNELS", 1))
    H = int(H if H is not None else params.INPUT_HEIGHT // params.R)
    W = int(W if W is not None else params.INPUT_WIDTH // params.R)
    L = Layout(H=H, W=W, hm=hm, nb_classes=int(params.NB_CLASSES), Cp=0, Ct=0,
               focal_a=float(params.FOCAL_LOSS_ALPHA), focal_b=float(params.FOCAL_LOSS_BETA),
               R=float(params.R), alpha=float(params.VARIANCE_ALPHA))
    idx = hm
    for key, f in params.REGRESSION_FIELDS.items():
        if not f.active:
            continue
        if key == "class":
            L.off_class = idx
        elif key == "r_offset":
            L.off_roff = idx
        elif key == "fullbox":
            L.off_box = idx
        elif key == "track_offset":
            L.off_track = idx
        if key == "3d_info":      # split like loss.py:24-29: radial (mape), orientation (mae + post), dims (mse)
            w = list(f.loss_weight)
            L.fields.append(("radial_dist", idx, 1, _lib.KIND_MAPE, _lib.POST_NONE, float(w[0])))
            L.fields.append(("orientation", idx + 1, 1, _lib.KIND_MAE, _lib.POST_ORIENT, float(w[1])))
            L.fields.append(("obj_dims", idx + 2, 3, _lib.KIND_MSE, _lib.POST_NONE, float(w[2])))
        else:
            L.fields.append((key, idx, int(f.size), _FIELD_KINDS[key], _lib.POST_NONE, float(f.loss_weight)))
        idx += int(f.size)
    L.Cp = idx
    L.Ct = idx + 1
    return L
