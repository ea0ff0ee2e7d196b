"""Feature interaction layers (reference: layers/feature_interaction/__init__.py:17-19).  Cross is on the hot path;
DotInteraction and MultiLayerDCN are the SURVEY 8f-4 widening."""
from .dcn import Cross
from .dot_interaction import DotInteraction
from .multi_layer_dcn import MultiLayerDCN
