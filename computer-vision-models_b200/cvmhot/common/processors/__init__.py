"""Pre-processor interface of the data generators, mirrored so that `ProcessImages` / `CenterTrackerProcess` plug into the
reference's `BaseDataGenerator` unchanged."""
from . import i_processor as _i

IPreProcessor = _i.IPreProcessor

__all__ = ["IPreProcessor"]
