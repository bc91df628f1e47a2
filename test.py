"""Hot-path subset of the reference's test.py: get_ensemble_weight (:25-50) and predict_location (:52-79).
The evaluation drivers of the reference (dataset loops, COCO export) are out of scope."""
from tracknetv3_b200.decode import predict_location, decode_heatmaps  # noqa: F401
from tracknetv3_b200.ensemble import get_ensemble_weight, TemporalEnsemble  # noqa: F401
