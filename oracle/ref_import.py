"""TEST INFRASTRUCTURE ONLY — loader for the *real* reference functions.

Imports the unmodified reference modules from /root/reference with the packages that are
absent from this image (tensorflow, albumentations, pymongo, matplotlib, ...) replaced by
MagicMock stubs, so that the NumPy/numba parts of the hot path run exactly as shipped:

  fill_heatmap            models/centernet/processor.py:17-38
  ProcessImages           models/centernet/processor.py:41  (calc_img_data :58-67, clip_to_img :46-56)
  process_2d_output       models/centernet/post_processing.py:6-66
  to_3channel             common/utils/image.py:72-100
  Roi, convert_back_to_roi common/utils/image.py:9-28
  CenternetParams         models/centernet/params.py:10
  CentertrackerParams     models/centertracker/params.py:3
  CenternetLoss           models/centernet/loss.py:6-155      (executed over oracle/tf_shim.py: TF ops on torch-CPU fp32)
  CentertrackerLoss       models/centertracker/loss.py:7-28   (same)
  MultitaskLoss           models/multitask/loss.py:13-47      (same; only calc_centernet is used)

The modules come from /root/reference in the build container and from its git-ignored copy oracle/_ref
(oracle/make_ref.py) on the GPU box.  Used by tests/golden/make_golden.py to generate the committed
fixtures, by the live-reference CPU tests and by bench.py's reference arm / cpu_baseline.  Nothing in the
product path imports this file.
"""
import os
import sys
from unittest.mock import MagicMock

# the mounted reference in the build container; its copy under oracle/_ref (oracle/make_ref.py) on the GPU box
_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("CVM_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "centernet")):
    REFERENCE_ROOT = os.path.join(_HERE, "_ref")

_STUBS = [
    "tensorflow", "tensorflow.keras", "tensorflow.keras.utils", "tensorflow.keras.losses",
    "tensorflow.keras.layers", "tensorflow.keras.models", "tensorflow.keras.regularizers",
    "tensorflow.keras.initializers", "tensorflow.keras.callbacks", "tensorflow.keras.optimizers",
    "tensorflow.python", "tensorflow.python.keras", "tensorflow.python.keras.utils",
    "tensorflow.python.keras.engine", "tensorflow.python.eager",
    "matplotlib", "matplotlib.pyplot", "albumentations", "pymongo", "pymongo.database",
    "pymongo.collection", "pygame", "redis", "tensorflow_model_optimization",
    "tflite_runtime", "tflite_runtime.interpreter", "pycoral", "pycoral.utils", "segmentation_models",
    "tensorflow.keras.applications", "tensorflow.keras.backend", "tensorflow.keras.metrics",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "centernet"))


_cache = {}


def load():
    """Return a dict of the real reference symbols (raises RuntimeError if not mounted)."""
    if _cache:
        return _cache
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    from . import tf_shim
    tf_shim.install()            # tensorflow, tensorflow.math/nn/keras/keras.losses: real semantics over torch fp32
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    # our own repo also has packages called `models`/`common` *inside* the product package, never at
    # top level, so putting the reference root first on sys.path is unambiguous.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from models.centernet.params import CenternetParams
    from models.centernet.processor import fill_heatmap, ProcessImages
    from models.centernet.post_processing import process_2d_output
    from models.centertracker.params import CentertrackerParams
    from common.utils.image import to_3channel, Roi, convert_back_to_roi
    from models.centernet.loss import CenternetLoss
    from models.centertracker.loss import CentertrackerLoss
    _cache.update(dict(
        CenternetParams=CenternetParams, CentertrackerParams=CentertrackerParams,
        fill_heatmap=fill_heatmap, ProcessImages=ProcessImages,
        process_2d_output=process_2d_output, to_3channel=to_3channel,
        Roi=Roi, convert_back_to_roi=convert_back_to_roi,
        CenternetLoss=CenternetLoss, CentertrackerLoss=CentertrackerLoss,
    ))
    try:        # the multitask package pulls in more of the reference (semseg/depth params, label spec); optional
        from models.multitask.loss import MultitaskLoss
        from models.multitask.params import MultitaskParams
        _cache.update(MultitaskLoss=MultitaskLoss, MultitaskParams=MultitaskParams)
    except Exception as e:      # pragma: no cover
        _cache["MultitaskLoss_error"] = repr(e)
    return _cache
