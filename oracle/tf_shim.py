"""TEST INFRASTRUCTURE ONLY — a minimal `tensorflow` stand-in over torch-CPU fp32.

TensorFlow is not installed in this image, so the reference's loss files cannot be executed as shipped.  This module
implements exactly the TF calls those files make, with TF's documented semantics, on torch CPU tensors in fp32, so
that the UNMODIFIED source files

    /root/reference/models/centernet/loss.py        (CenternetLoss: obj_focal_loss, calc_loss, call, every wrapper)
    /root/reference/models/centertracker/loss.py    (CentertrackerLoss)
    /root/reference/models/multitask/loss.py        (MultitaskLoss.calc_centernet's slicing)

can be imported and RUN line by line (oracle/ref_import.py installs it in sys.modules before the import).
tests/golden/make_golden.py uses that to emit tests/golden/loss_*.npz; oracle/loss_np.py and the CUDA path are then
checked against vectors produced by executing reference lines loss.py:31-155, not by a restatement of them.

What is restated here is TensorFlow's op semantics (each function cites the TF API it stands for), not the reference:
  * every op computes in the dtype of its operands (fp32 after the tf.cast in loss.py:134-135); Python scalars are
    weak-typed like in TF; reductions run over all elements in fp32 (TF/Eigen and torch use different summation
    trees, both with errors far below the 1e-5 parity bar at the fixture sizes);
  * tf.equal / tf.less / tf.greater return bool tensors, tf.cast(bool, float32) gives 0/1;
  * tf.math.pow = elementwise pow, tf.clip_by_value = clamp (NaN propagates);
  * tf.math.multiply_no_nan(x, y) = 0 where y == 0 (even if x is NaN/inf), else x*y;
  * tf.cond(pred, a, b) in eager mode calls exactly one branch;
  * tf.keras.losses.categorical_crossentropy(y_true, y_pred, from_logits=True) =
    tf.nn.softmax_cross_entropy_with_logits(labels, logits, axis=-1) = -sum_c labels_c * log_softmax(logits)_c
    (labels are NOT renormalised on the logits path);
  * keras.losses.Loss.__call__(y_true, y_pred) = call() followed by the AUTO reduction, the identity for the 0-d
    tensor these losses return.
Unknown attributes fall back to MagicMock, so the model/callback modules that the reference packages import next to
the loss still import (they are never executed).

Nothing in the product path imports this file.
"""
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
bool_ = torch.bool


def _t(x, like=None):
    """tf.convert_to_tensor: numpy / python -> tensor; Python scalars stay weak-typed (handled by torch promotion)."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x))
    return x   # python scalar: torch's own weak scalar promotion matches TF's here (fp32 tensor op scalar -> fp32)


def cast(x, dtype):                                   # tf.cast
    x = _t(x)
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(x)
    return x.to(dtype)


def equal(a, b):                                      # tf.equal
    return torch.eq(_t(a), b)


def less(a, b):                                       # tf.less
    return torch.lt(_t(a), b)


def greater(a, b):                                    # tf.greater
    return torch.gt(_t(a), b)


def clip_by_value(x, lo, hi):                         # tf.clip_by_value
    return torch.clamp(_t(x), min=lo, max=hi)


def reduce_sum(x, axis=None, keepdims=False):         # tf.reduce_sum
    x = _t(x)
    if axis is None:
        return x.sum()
    return x.sum(dim=axis, keepdim=keepdims)


def reduce_max(x, axis=None, keepdims=False):         # tf.reduce_max
    x = _t(x)
    if axis is None:
        return x.max()
    return x.max(dim=axis, keepdim=keepdims).values


def stack(values, axis=0):                            # tf.stack
    return torch.stack([_t(v) for v in values], dim=axis)


def shape(x):                                         # tf.shape (only consumed by broadcast_to)
    return tuple(_t(x).shape)


def broadcast_to(x, shp):                             # tf.broadcast_to
    return torch.broadcast_to(_t(x), tuple(int(s) for s in shp))


def maximum(a, b):                                    # tf.maximum
    a = _t(a)
    if not isinstance(b, torch.Tensor):
        b = torch.tensor(b, dtype=a.dtype)
    return torch.maximum(a, b)


def squeeze(x, axis=None):                            # tf.squeeze
    return _t(x).squeeze() if axis is None else _t(x).squeeze(axis)


def cond(pred, true_fn, false_fn):                    # tf.cond, eager: exactly one branch runs
    return true_fn() if bool(pred) else false_fn()


def convert_to_tensor(x, dtype=None):
    x = _t(x)
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(x)
    return x if dtype is None else x.to(dtype)


def _multiply_no_nan(x, y):                           # tf.math.multiply_no_nan
    x, y = _t(x), _t(y)
    return torch.where(y == 0, torch.zeros((), dtype=x.dtype), x * y)


def _squared_difference(a, b):                        # tf.math.squared_difference
    d = _t(a) - _t(b)
    return d * d


def _pow(x, y):                                       # tf.math.pow
    return torch.pow(_t(x), y)


def _softmax_cross_entropy_with_logits(labels, logits, axis=-1):   # tf.nn.softmax_cross_entropy_with_logits
    labels, logits = _t(labels), _t(logits)
    return -(labels * torch.log_softmax(logits, dim=axis)).sum(dim=axis)


def _categorical_crossentropy(y_true, y_pred, from_logits=False, label_smoothing=0.0, axis=-1):
    """tf.keras.losses.categorical_crossentropy (keras/backend.py: categorical_crossentropy)."""
    y_true, y_pred = _t(y_true), _t(y_pred)
    assert label_smoothing == 0.0
    if from_logits:
        return _softmax_cross_entropy_with_logits(y_true, y_pred, axis)
    # probability path: renormalise, clip to [eps, 1-eps], -sum t*log p  (not used by the CenterNet loss)
    eps = 1e-7
    p = y_pred / y_pred.sum(dim=axis, keepdim=True)
    p = torch.clamp(p, eps, 1.0 - eps)
    return -(y_true * torch.log(p)).sum(dim=axis)


class Loss:
    """tf.keras.losses.Loss: __call__ = call + AUTO reduction (identity on the scalars these losses return)."""

    def __init__(self, reduction="auto", name=None):
        self.reduction = reduction
        self.name = name

    def call(self, y_true, y_pred):
        raise NotImplementedError

    def __call__(self, y_true, y_pred, sample_weight=None):
        assert sample_weight is None
        out = self.call(_t(y_true), _t(y_pred))
        out = _t(out)
        return out if out.dim() == 0 else out.mean()      # SUM_OVER_BATCH_SIZE on a non-scalar


# ---- eager-mode stand-ins for the glue calls of cvmhot/keras_shim.py (tests/test_gpu_keras_shim.py): in eager TF,
#      tf.py_function runs the Python function at once, tf.custom_gradient returns the forward value and keeps the gradient
#      function for the tape, and the DLPack hand-over is zero-copy.  Here the "tape" is an attribute on the value.
def py_function(func, inp, Tout=None):                # tf.py_function, eager
    out = func(*inp)
    if isinstance(out, torch.Tensor) and not hasattr(out, "set_shape"):
        pass
    return out


def custom_gradient(f):                               # tf.custom_gradient, eager: value now, grad_fn kept with it
    def wrapped(*args):
        value, grad_fn = f(*args)
        value._grad_fn_for_tape = grad_fn
        return value
    return wrapped


def reshape(x, shp):                                  # tf.reshape
    return _t(x).reshape(tuple(int(v) for v in shp))


def _dl_to(t):                                        # tf.experimental.dlpack.to_dlpack: a torch tensor is its own capsule
    return t


def _dl_from(capsule):                                # tf.experimental.dlpack.from_dlpack
    return torch.utils.dlpack.from_dlpack(capsule)


class _Shim(types.ModuleType):
    """A module whose unknown attributes are MagicMocks (so unrelated `from tensorflow.x import y` lines still import)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


def build_modules():
    """-> {module name: module} for sys.modules."""
    tf = _Shim("tensorflow")
    for k, v in dict(float32=float32, float64=float64, int32=int32, bool=bool_, cast=cast, equal=equal, less=less,
                     greater=greater, clip_by_value=clip_by_value, reduce_sum=reduce_sum, reduce_max=reduce_max,
                     stack=stack, shape=shape, broadcast_to=broadcast_to, maximum=maximum, squeeze=squeeze, cond=cond,
                     convert_to_tensor=convert_to_tensor, Tensor=torch.Tensor).items():
        setattr(tf, k, v)
    m = _Shim("tensorflow.math")
    for k, v in dict(pow=_pow, log=lambda x: torch.log(_t(x)), abs=lambda x: torch.abs(_t(x)),
                     sqrt=lambda x: torch.sqrt(_t(x)), cos=lambda x: torch.cos(_t(x)),
                     squared_difference=_squared_difference, multiply_no_nan=_multiply_no_nan,
                     reduce_sum=reduce_sum, reduce_max=reduce_max, maximum=maximum).items():
        setattr(m, k, v)
    tf.math = m
    nn = _Shim("tensorflow.nn")
    nn.softmax_cross_entropy_with_logits = _softmax_cross_entropy_with_logits
    tf.nn = nn
    tf.py_function = py_function
    tf.custom_gradient = custom_gradient
    tf.reshape = reshape
    exp_ = _Shim("tensorflow.experimental")
    dl = _Shim("tensorflow.experimental.dlpack")
    dl.to_dlpack = _dl_to
    dl.from_dlpack = _dl_from
    exp_.dlpack = dl
    tf.experimental = exp_
    if not hasattr(torch.Tensor, "set_shape"):        # EagerTensor.set_shape: a static-shape hint, nothing to do here
        torch.Tensor.set_shape = lambda self, shape: None
    keras = _Shim("tensorflow.keras")
    losses = _Shim("tensorflow.keras.losses")
    losses.Loss = Loss
    losses.categorical_crossentropy = _categorical_crossentropy
    keras.losses = losses
    tf.keras = keras
    return {"tensorflow": tf, "tensorflow.math": m, "tensorflow.nn": nn, "tensorflow.keras": keras,
            "tensorflow.keras.losses": losses}


# torch tensors already have .numpy(), slicing, .shape and arithmetic with Python floats in fp32, which is all the
# reference's loss code asks of an EagerTensor.


def install():
    """Put the shim in sys.modules (idempotent).  Returns the `tensorflow` module object."""
    cur = sys.modules.get("tensorflow")
    if isinstance(cur, _Shim):
        return cur
    mods = build_modules()
    sys.modules.update(mods)
    return mods["tensorflow"]
