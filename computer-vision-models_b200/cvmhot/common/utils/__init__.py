from .image import Roi, convert_back_to_roi, to_3channel
