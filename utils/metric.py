"""Drop-in for the reference's utils/metric.py (WBCELoss :3-20, get_metric :22-46)."""
from tracknetv3_b200.metric import WBCELoss, get_metric  # noqa: F401
