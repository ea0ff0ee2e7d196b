"""Feature interaction layers (reference: layers/feature_interaction/__init__.py:17-19; only Cross is on
the hot path -- DotInteraction / MultiLayerDCN are out of scope)."""
from .dcn import Cross
