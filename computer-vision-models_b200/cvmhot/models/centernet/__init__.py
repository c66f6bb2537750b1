from .params import CenternetParams
from .processor import ProcessImages, fill_heatmap
from .loss import CenternetLoss
from .post_processing import process_2d_output, decode_topk
