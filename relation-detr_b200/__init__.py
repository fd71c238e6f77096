"""relation-detr_b200: B200-native (sm_100a) replacement for the Relation-DETR hot path.

Behind the reference's own module API:

* ``MultiScaleDeformableAttention`` forward/backward  (reference: models/bricks/ms_deform_attn.py)
* ``PositionRelationEmbedding``                        (reference: models/bricks/relation_transformer.py:481-532)
* ``HungarianMatcher`` (row N3): cost matrices + assignment on the device, SciPy-identical pairs
                                                        (reference: models/matcher/hungarian_matcher.py)

* ``graphs.capture_static_parts``: CUDA-graph capture of the static-shape parts of the reference's detector around these operators

The kernels live in ``csrc/`` and are reached through the C-ABI library ``librdetr_ops.so``
(declared in ``include/rdetr_ops.h``).  There is no CPU fallback: calling an operator without the
library, or with non-CUDA tensors, raises.
"""
__version__ = "0.1.0"

from . import graphs, ops, workloads  # noqa: E402,F401
from .matcher import HungarianMatcher  # noqa: E402,F401
from .modules import MultiScaleDeformableAttention, PositionRelationEmbedding  # noqa: E402,F401
from .ops import (MultiScaleDeformableAttnFunction, ms_deform_attn, position_relation_bias,  # noqa: E402,F401
                  relation_dim_t)
