"""TEST INFRASTRUCTURE ONLY — fp64 NumPy restatement of the CenterNet / CenterTracker loss.

TensorFlow is not installed in this image, so the TF loss cannot be executed; this file restates it
op by op from the source:
  CenternetLoss.obj_focal_loss   /root/reference/models/centernet/loss.py:31-60
  CenternetLoss.calc_loss        loss.py:107-131
  CenternetLoss.call             loss.py:133-155
  CentertrackerLoss              /root/reference/models/centertracker/loss.py:16-28
  MultitaskLoss.calc_centernet   /root/reference/models/multitask/loss.py:21-24,44-47

PINNED TO THE REFERENCE SOURCE (round 2): oracle/tf_shim.py supplies TensorFlow's ops over torch-CPU fp32, so the
UNMODIFIED reference files run line by line (oracle/ref_import.py); tests/golden/make_golden.py stores their outputs in
tests/golden/loss_*.npz (the reference's own fixtures of loss_test.py:9-49 and centertracker/loss_test.py:10-21 with the
perturbations its tests apply, Profile R / N at 128x384, tracker layouts, all fields, no objects, clip edges, the
multitask slice) and tests/test_oracle_loss_golden.py holds every term of this restatement to them within 1e-5
relative.  (The reference's own tests carry no numeric expectations: inequalities only, several stale against the
current class-CE loss; the still-valid ones and the derived values 0.27574 / 0.28467 / 0.28311 /
4.88091 are asserted in tests/test_oracle_loss.py, and an independently written torch fp32 twin, `loss_torch32`, must
agree to 1e-5.)
TF semantics honoured: tf.equal/tf.less on fp32 values; pow on the UNclipped prediction; clip only inside
log; categorical_crossentropy(from_logits=True) = -sum_c t_c*log_softmax(p)_c with labels not
renormalised; multiply_no_nan; tf.cond(n>0, ..); every reduction runs over all axes including batch.
"""
import numpy as np

from .layout import Layout, KIND_MSE, KIND_MAE, KIND_MAPE, KIND_CE, POST_ORIENT


def focal_sums(L: Layout, y_true, y_pred, use_weights=True):
    """loss.py:32-57 -> (P, N, n). y_true [B,H,W,Ct] (weights plane last), y_pred [B,H,W,Cp]."""
    Y = y_true[..., :L.hm].astype(np.float32)
    Yh32 = y_pred[..., :L.hm].astype(np.float32)
    pos = (Y == np.float32(1.0)).astype(np.float64)                         # :35 (tf.cast of the mask)
    neg = (Y < np.float32(1.0)).astype(np.float64)                          # :36
    Y = Y.astype(np.float64)
    Yh = Yh32.astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        pl = -pos * np.power(1.0 - Yh, L.focal_a) * np.log(np.clip(Yh, 0.01, 0.99))               # :38-42
        nl = (-neg * np.power(1.0 - Y, L.focal_b) * np.power(Yh, L.focal_a)
              * np.log(np.clip(1.0 - Yh, 0.01, 0.99)))                                            # :43-48
    n = float(pos.sum())                                                    # :50
    if use_weights:
        w = y_true[..., -1].astype(np.float64)[..., None]                   # :137, :52
        P = float((pl * w).sum())
        N = float((nl * w).sum())                                           # :53-54
    else:
        P = float(pl.sum())
        N = float(nl.sum())                                                 # :56-57
    return P, N, n


def focal(L, y_true, y_pred, use_weights=True):
    P, N, n = focal_sums(L, y_true, y_pred, use_weights)
    return (P + N) / n if n > 0 else N                                      # :59


def pos_mask(L, y_true):
    return (y_true[..., :L.hm].astype(np.float32) == np.float32(1.0)).any(axis=-1)   # :110-111


def field_sum(L, f, y_true, y_pred):
    """loss.py:115-128: sum of |masked term| for one regression field (before the /nb_objects)."""
    pm = pos_mask(L, y_true)
    t = y_true[..., f.off:f.off + f.size].astype(np.float64)[pm]
    p = y_pred[..., f.off:f.off + f.size].astype(np.float64)[pm]
    if f.kind == KIND_MSE:
        m = (t - p) ** 2                                                    # :116
    elif f.kind == KIND_MAE:
        m = (t - p)                                                         # :118
    elif f.kind == KIND_MAPE:
        m = (t - p) / np.maximum(np.abs(t), 1.0)                            # :120
    elif f.kind == KIND_CE:
        mx = p.max(axis=-1, keepdims=True)
        lse = mx + np.log(np.exp(p - mx).sum(axis=-1, keepdims=True))
        m = -(t * (p - lse)).sum(axis=-1)                                   # :122 (labels not renormalised)
    else:
        raise AssertionError                                                # :126
    return float(np.abs(m).sum())                                           # :127-128


def field_loss(L, f, y_true, y_pred):
    nobj = float(pos_mask(L, y_true).sum())                                 # :113
    s = field_sum(L, f, y_true, y_pred)
    v = s / nobj if nobj > 0 else s                                         # :130
    if f.post == POST_ORIENT:
        v = np.sqrt(1.0 - 0.99 * np.cos(2.0 * v)) + abs(v * v * 0.05) - 0.0999     # :98
    return float(v)


def partials(L, y_true, y_pred, use_weights=True):
    """The vector that is all-reduced across GPUs: [P, N, n, nobj, field sums...]."""
    P, N, n = focal_sums(L, y_true, y_pred, use_weights)
    nobj = float(pos_mask(L, y_true).sum())
    return np.array([P, N, n, nobj] + [field_sum(L, f, y_true, y_pred) for f in L.fields], dtype=np.float64)


def finalize(L, part):
    """loss.py:59,130,140-153 applied to (already summed) partials -> (total, [focal, field...])."""
    P, N, n, nobj = part[:4]
    terms = [(P + N) / n if n > 0 else N]
    total = terms[0]
    for i, f in enumerate(L.fields):
        s = part[4 + i]
        v = s / nobj if nobj > 0 else s
        if f.post == POST_ORIENT:
            v = np.sqrt(1.0 - 0.99 * np.cos(2.0 * v)) + abs(v * v * 0.05) - 0.0999
        terms.append(float(v))
        total += v * f.weight
    return float(total), terms


def total_loss(L, y_true, y_pred):
    """CenternetLoss.call / CentertrackerLoss.call (loss.py:133-155, centertracker/loss.py:22-28)."""
    return finalize(L, partials(L, y_true, y_pred, True))


def loss_torch32(L, y_true, y_pred):
    """Independent fp32 torch-CPU twin of `total_loss`, written against the TF source rather than against
    the NumPy code above (different op order: mask-multiply like TF instead of boolean indexing)."""
    import torch
    yt = torch.as_tensor(np.ascontiguousarray(y_true), dtype=torch.float32)
    yp = torch.as_tensor(np.ascontiguousarray(y_pred), dtype=torch.float32)
    w = yt[..., -1]
    yt = yt[..., :-1]
    Y, Yh = yt[..., :L.hm], yp[..., :L.hm]
    pos = (Y == 1.0).float()
    neg = (Y < 1.0).float()
    pl = -pos * torch.pow(1.0 - Yh, L.focal_a) * torch.log(torch.clamp(Yh, 0.01, 0.99))
    nl = -neg * torch.pow(1.0 - Y, L.focal_b) * torch.pow(Yh, L.focal_a) * torch.log(torch.clamp(1.0 - Yh, 0.01, 0.99))
    n = pos.sum()
    sw = torch.stack([w] * L.hm, dim=-1)
    P, N = (pl * sw).double().sum(), (nl * sw).double().sum()
    total = (P + N) / n if n > 0 else N
    pm = pos.max(dim=-1, keepdim=True).values
    nobj = pm.sum()
    for f in L.fields:
        t, p = yt[..., f.off:f.off + f.size], yp[..., f.off:f.off + f.size]
        if f.kind == KIND_MSE:
            m = pm * (t - p) ** 2
        elif f.kind == KIND_MAE:
            m = pm * (t - p)
        elif f.kind == KIND_MAPE:
            m = pm * ((t - p) / torch.clamp(t.abs(), min=1.0))
        else:
            ce = -(t * torch.log_softmax(p, dim=-1)).sum(dim=-1)
            m = torch.where(pm[..., 0] > 0, ce, torch.zeros_like(ce))
        s = m.abs().double().sum()
        v = s / nobj if nobj > 0 else s
        if f.post == POST_ORIENT:
            v = torch.sqrt(1.0 - 0.99 * torch.cos(2.0 * v)) + (v * v * 0.05).abs() - 0.0999
        total = total + v * f.weight
    return float(total)
