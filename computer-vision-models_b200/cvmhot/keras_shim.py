"""TensorFlow / Keras front for the CUDA loss: `model.compile(loss=CenternetLoss(params), metrics=[loss.class_loss, ...])`
keeps working literally (reference models/centernet/loss.py:6, models/centernet/train.py:62).

NOT EXERCISED AGAINST REAL TENSORFLOW: it is not installed in the build image, so this module is
the binding a maintainer of the reference adds on a box that has TF.  tests/test_gpu_keras_shim.py drives its glue (DLPack
hand-over, py_function bodies, the custom-gradient pair, every metric method) through the eager TensorFlow stand-in that
the test suite owns and checks values and gradients against the executed reference; everything below the DLPack hand-over
is the tested torch-facing mirror (cvmhot.models.centernet.loss).  Tensors stay on the GPU: TF -> DLPack -> torch (zero copy),
libcvmhot kernels on torch's current stream, result -> DLPack -> TF.  The gradient goes through tf.custom_gradient to the
hand-written backward kernel (cvm_loss_bwd).

    from cvmhot.keras_shim import CenternetLoss          # instead of models.centernet.loss
    loss = CenternetLoss(params)
    model.compile(optimizer=opt, loss=loss, metrics=[loss.class_loss, loss.r_offset_loss, loss.fullbox_loss])
"""
import torch
from torch.utils import dlpack as _tdl

try:
    import tensorflow as tf
except ImportError as e:      # pragma: no cover - the whole point of the module
    raise ImportError("cvmhot.keras_shim needs TensorFlow; the torch-facing mirror is cvmhot.models.centernet.loss") from e

from cvmhot.models.centernet.loss import CenternetLoss as _TorchLoss
from cvmhot.models.centertracker.loss import CentertrackerLoss as _TorchTrackerLoss


def _to_torch(t):
    return _tdl.from_dlpack(tf.experimental.dlpack.to_dlpack(t))


def _to_tf(t):
    return tf.experimental.dlpack.from_dlpack(_tdl.to_dlpack(t.contiguous()))


class CenternetLoss(tf.keras.losses.Loss):
    _torch_cls = _TorchLoss

    def __init__(self, params, **kw):
        super().__init__(**kw)
        self.params = params
        self._impl = self._torch_cls(params)

    # ---- total loss with the hand-written backward --------------------------------------------------------------------
    def call(self, y_true, y_pred):
        y_true = tf.cast(y_true, tf.float32)
        y_pred = tf.cast(y_pred, tf.float32)

        @tf.custom_gradient
        def _loss(y_pred_):
            def fwd(yt, yp):
                ypt = _to_torch(yp).requires_grad_(True)
                out = self._impl.call(_to_torch(yt), ypt)
                torch.cuda.current_stream().synchronize()      # TF and torch use different streams
                self._saved = (out, ypt)
                return _to_tf(out.detach().reshape(1))[0]

            value = tf.py_function(fwd, [y_true, y_pred_], tf.float32)
            value.set_shape(())

            def grad(upstream):
                def bwd(up):
                    out, ypt = self._saved
                    (g,) = torch.autograd.grad(out, ypt, grad_outputs=_to_torch(tf.reshape(up, [1]))[0],
                                               retain_graph=True)      # (a persistent tape may ask twice)
                    torch.cuda.current_stream().synchronize()
                    return _to_tf(g)
                g = tf.py_function(bwd, [upstream], tf.float32)
                g.set_shape(y_pred_.shape)
                return g
            return value, grad
        return _loss(y_pred)

    # ---- the sub-terms Keras uses as metrics: one fused pass for all of them (the mirror caches it per batch) --------------
    def _metric(self, name, y_true, y_pred, *extra):
        def run(yt, yp):
            out = getattr(self._impl, name)(_to_torch(tf.cast(yt, tf.float32)), _to_torch(tf.cast(yp, tf.float32)), *extra)
            torch.cuda.current_stream().synchronize()
            return _to_tf(out.reshape(1))[0]
        v = tf.py_function(run, [y_true, y_pred], tf.float32)
        v.set_shape(())
        return v

    def obj_focal_loss(self, y_true, y_pred, weights=None):
        if weights is not None:       # (only `call` passes weights, loss.py:137-140, and call() is fused here)
            raise NotImplementedError("the weighted focal term is part of call(); as a metric obj_focal_loss takes no weights")
        return self._metric("obj_focal_loss", y_true, y_pred)

    def class_loss(self, y_true, y_pred):
        return self._metric("class_loss", y_true, y_pred)

    def r_offset_loss(self, y_true, y_pred):
        return self._metric("r_offset_loss", y_true, y_pred)

    def fullbox_loss(self, y_true, y_pred):
        return self._metric("fullbox_loss", y_true, y_pred)

    def l_shape_loss(self, y_true, y_pred):
        return self._metric("l_shape_loss", y_true, y_pred)

    def radial_dist_loss(self, y_true, y_pred):
        return self._metric("radial_dist_loss", y_true, y_pred)

    def orientation_loss(self, y_true, y_pred):
        return self._metric("orientation_loss", y_true, y_pred)

    def obj_dims_loss(self, y_true, y_pred):
        return self._metric("obj_dims_loss", y_true, y_pred)


class CentertrackerLoss(CenternetLoss):
    _torch_cls = _TorchTrackerLoss

    def track_offset_loss(self, y_true, y_pred):
        return self._metric("track_offset_loss", y_true, y_pred)
