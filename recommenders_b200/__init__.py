"""recommenders_b200 -- a B200-native (sm_100a) implementation of the TensorFlow Recommenders retrieval /
ranking hot path behind the reference's own API surface:

    import recommenders_b200 as tfrs
    tfrs.Model, tfrs.tasks.Retrieval, tfrs.metrics.FactorizedTopK,
    tfrs.layers.factorized_top_k.{BruteForce, Streaming}, tfrs.layers.dcn.Cross

(namespace per tensorflow_recommenders/__init__.py:51-61 and layers/__init__.py:18-23).  Tensors are CUDA
torch tensors; all arithmetic on the path runs in libtfrs_b200.so (include/tfrs_b200.h).  No CPU fallback.
"""
from . import data
from . import layers
from . import metrics
from . import models
from . import optimizers
from . import tasks
from .models import Model

__version__ = "0.1.0"
