"""Drop-in for the reference's model.py: same class names and signatures (reference model.py:4-129),
implemented by hand-written sm_100a kernels in tracknetv3_b200."""
from tracknetv3_b200.model import (Conv2DBlock, Double2DConv, Triple2DConv, TrackNet,  # noqa: F401
                                   Conv1DBlock, Double1DConv, InpaintNet)
