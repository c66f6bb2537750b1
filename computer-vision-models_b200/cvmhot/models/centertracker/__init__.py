from .params import CentertrackerParams
from .processor import CenterTrackerProcess
from .loss import CentertrackerLoss
from .tracking import Tracker
