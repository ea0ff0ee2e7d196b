"""Layers namespace, shaped like tensorflow_recommenders/layers/__init__.py:18-23."""
from . import embedding
from . import factorized_top_k
from . import feature_interaction
from . import loss
from .feature_interaction import dcn
