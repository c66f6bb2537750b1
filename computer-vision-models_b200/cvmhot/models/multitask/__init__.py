from .params import MultitaskParams
from .loss import MultitaskLoss
