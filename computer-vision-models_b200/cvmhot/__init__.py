"""cvmhot — B200-native CenterNet/CenterTracker heatmap hot path (render + loss + decode).

Python host side of libcvmhot.so.  Layout of this package mirrors the reference repository for the files on the
hot path, so `from cvmhot.models.centernet import CenternetParams, ProcessImages, CenternetLoss, process_2d_output`
reads like the reference's `from models.centernet import ...` (reference models/centernet/__init__.py:1-7).
"""
from . import _lib
from .layout import Layout, layout_from_params

__all__ = ["_lib", "Layout", "layout_from_params"]
