"""dream.spatial_softmax -> dream_b200.spatial_softmax (SoftArgmaxPavlo, dream/spatial_softmax.py)."""
from dream_b200.spatial_softmax import SoftArgmaxPavlo   # noqa: F401
