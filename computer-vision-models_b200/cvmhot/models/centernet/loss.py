"""CenterNet loss on the GPU: one fused streaming pass instead of ~60 eager ops.

Same public surface as the reference's CenternetLoss (models/centernet/loss.py:6-155): `loss(y_true, y_pred)` /
`.call(...)` return the scalar total; the sub-terms used as Keras metrics (`obj_focal_loss`, `class_loss`,
`r_offset_loss`, `fullbox_loss`, `l_shape_loss`, `radial_dist_loss`, `orientation_loss`, `obj_dims_loss`, `calc_loss`)
stay individually callable.  Tensors are torch CUDA tensors in the reference's NHWC layout.

Data parallel: the loss is normalised by BATCH-GLOBAL counts (loss.py:50,113), so with the batch sharded over GPUs the
partial sums/counts are all-reduced (one 128-byte NCCL all-reduce) before the finalise step — pass `process_group`.
"""
import torch

from cvmhot import _lib, ops
from cvmhot.layout import layout_from_params


def _distributed(group):
    return group is not None and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1


def _loss_vector(layout, y_true, y_pred, use_weights, group):
    """-> (out [total, focal, fields...], partials).  One launch on a single device (cvm_loss_fwd_total); with the batch
    sharded over GPUs: partials, the ONE exchange of the path (an all-gather of 128 bytes per rank), then the rank-ordered
    sum and the finalise step in one launch - bit-identical on every rank and to a single GPU working through the shards."""
    if not _distributed(group):
        return ops.loss_total(layout, y_true, y_pred, use_weights)
    partials = ops.loss_partials(layout, y_true, y_pred, use_weights)
    world = torch.distributed.get_world_size(group)
    gathered = torch.empty((world, _lib.CVM_NPART), dtype=torch.float64, device=partials.device)
    torch.distributed.all_gather_into_tensor(gathered.view(-1), partials, group=group)
    return ops.loss_finalize_gathered(layout, gathered)


class _LossFn(torch.autograd.Function):
    """total loss with a hand-written backward (cvm_loss_bwd)."""

    @staticmethod
    def forward(ctx, y_pred, y_true, layout, group):
        out, partials = _loss_vector(layout, y_true, y_pred, True, group)
        ctx.layout = layout
        ctx.save_for_backward(y_true, y_pred, partials)
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        y_true, y_pred, partials = ctx.saved_tensors
        up = grad_out.to(torch.float32).reshape(1).contiguous()
        grad = ops.loss_backward(ctx.layout, y_true, y_pred, partials, upstream=up)
        if grad.shape[-1] != y_pred.shape[-1]:
            pad = torch.zeros(y_pred.shape, dtype=grad.dtype, device=grad.device)
            pad[..., :grad.shape[-1]] = grad
            grad = pad
        return grad, None, None, None


class CenternetLoss:
    def __init__(self, params, process_group=None):
        self.params = params
        self.process_group = process_group
        self.obj_pos = [0, getattr(params, "HM_CHANNELS", 1)]
        self._cache = {}

    # ---- helpers -------------------------------------------------------------------------------------------------------
    def _layout(self, y_true):
        return layout_from_params(self.params, H=int(y_true.shape[-3]), W=int(y_true.shape[-2]))

    @staticmethod
    def _f32(t):
        return t if t.dtype == torch.float32 else t.to(torch.float32)       # tf.cast(..., tf.float32), loss.py:134-135

    def _vector(self, y_true, y_pred, use_weights):
        """The finalised loss vector of (y_true, y_pred), computed ONCE per pair of tensors: the reference's
        `metrics=[loss.class_loss, loss.r_offset_loss, loss.fullbox_loss, ...]` (train.py:62) calls every sub-term on the same
        batch, and each term falls out of the same streaming pass.  An entry is reused only for the same storage, view
        geometry and version counter (an in-place update or a new batch misses); it keeps its tensors alive, so their memory
        cannot be handed to another tensor while the entry exists (one entry per weighting mode: at most the last batch)."""
        L = self._layout(y_true)

        def sig(t):
            return (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape[:-1]), t.stride(-2), t._version)

        key = (sig(y_true), sig(y_pred))
        hit = self._cache.get(bool(use_weights))
        if hit is None or hit[0] != key:
            out, _ = _loss_vector(L, y_true, y_pred, use_weights, self.process_group)
            hit = self._cache[bool(use_weights)] = (key, out, L, y_true, y_pred)
        return hit[1], hit[2]

    def terms(self, y_true, y_pred, use_weights=True):
        """All terms from ONE pass: dict(total, focal, <field names...>) of 0-d CUDA tensors.  y_true carries the weights plane."""
        y_true, y_pred = self._f32(y_true), self._f32(y_pred)
        out, L = self._vector(y_true, y_pred, use_weights)
        d = {"total": out[0], "obj_focal": out[1]}
        for i, name in enumerate(L.field_names()):
            d[name] = out[2 + i]
        return d

    def _term(self, y_true, y_pred, name):
        # metric calls receive y_true with or without the weights plane; both are channel-prefix views of the same pixels
        y_true, y_pred = self._f32(y_true), self._f32(y_pred)
        L = self._layout(y_true)
        if y_true.shape[-1] < L.Cp:
            raise _lib.CvmError("y_true has fewer channels than the layout")
        out, L = self._vector(y_true, y_pred, False)
        return out[2 + L.field_names().index(name)]

    # ---- reference API -------------------------------------------------------------------------------------------------
    def __call__(self, y_true, y_pred):
        return self.call(y_true, y_pred)

    def call(self, y_true, y_pred):
        """total = focal + sum(field loss * loss_weight)  (reference loss.py:133-155); differentiable wrt y_pred."""
        y_true, y_pred = self._f32(torch.as_tensor(y_true)), self._f32(torch.as_tensor(y_pred))
        L = self._layout(y_true)
        if y_true.shape[-1] < L.Ct:
            raise _lib.CvmError(f"y_true needs {L.Ct} channels (weights plane last), got {y_true.shape[-1]}")
        return _LossFn.apply(y_pred, y_true, L, self.process_group)

    def obj_focal_loss(self, y_true, y_pred, weights=None):
        """Penalty-reduced pixel focal loss (reference loss.py:31-60).  `weights` must be None (metric mode) or the
        weights plane of the very tensor y_true is a view of (how `call` uses it, loss.py:137-140)."""
        y_true, y_pred = self._f32(y_true), self._f32(y_pred)
        L = self._layout(y_true)
        use_w = weights is not None
        if use_w:
            st = y_true.stride(-2)
            if weights.data_ptr() != y_true.data_ptr() + 4 * (L.Ct - 1) or weights.stride(-1) != st:
                raise _lib.CvmError("weights must be y_true[..., -1] of the full ground-truth tensor")
        return self._vector(y_true, y_pred, use_w)[0][1]

    def class_loss(self, y_true, y_pred):
        return self._term(y_true, y_pred, "class")

    def r_offset_loss(self, y_true, y_pred):
        return self._term(y_true, y_pred, "r_offset")

    def fullbox_loss(self, y_true, y_pred):
        return self._term(y_true, y_pred, "fullbox")

    def l_shape_loss(self, y_true, y_pred):
        return self._term(y_true, y_pred, "l_shape")

    def radial_dist_loss(self, y_true, y_pred):
        return self._term(y_true, y_pred, "radial_dist")

    def orientation_loss(self, y_true, y_pred):
        return self._term(y_true, y_pred, "orientation")

    def obj_dims_loss(self, y_true, y_pred):
        return self._term(y_true, y_pred, "obj_dims")

    def calc_loss(self, y_true, y_true_feat, y_pred_feat, loss_type: str = "mse"):
        """Generic masked regression term on arbitrary feature slices (reference loss.py:107-131)."""
        kinds = {"mse": _lib.KIND_MSE, "mae": _lib.KIND_MAE, "mape": _lib.KIND_MAPE, "cross_entropy": _lib.KIND_CE}
        if loss_type not in kinds:
            raise AssertionError(loss_type)                                  # reference :126
        hm = self.obj_pos[1]
        f = int(y_true_feat.shape[-1])
        # assemble a compact [hm | feature | weights] ground truth and [hm | feature] prediction (device-side glue only)
        yt = torch.cat([self._f32(y_true[..., :hm]), self._f32(y_true_feat),
                        torch.ones_like(y_true_feat[..., :1], dtype=torch.float32)], dim=-1).contiguous()
        yp = torch.cat([torch.zeros_like(yt[..., :hm]), self._f32(y_pred_feat)], dim=-1).contiguous()
        L = self._layout(y_true)
        L.Cp, L.Ct = hm + f, hm + f + 1
        L.off_class = L.off_roff = L.off_box = L.off_track = -1
        L.fields = [("feat", hm, f, kinds[loss_type], _lib.POST_NONE, 1.0)]
        return _loss_vector(L, yt, yp, False, self.process_group)[0][2]
