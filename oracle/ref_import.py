"""Import the two reference operators from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing that runs
there may call this; it is used by ``make_golden.py`` and by CPU tests that skip when the reference
tree is absent.  ``util.misc`` is replaced by a 4-line stand-in because the real module drags in
packages that are not installed here; the stand-in keeps the semantics of
``util/misc.py:31-35``.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("RDETR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "bricks", "ms_deform_attn.py"))


def load():
    """-> (multi_scale_deformable_attn_pytorch, PositionRelationEmbedding, box_rel_encoding, MultiScaleDeformableAttention)"""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "util.misc" not in sys.modules:
        util_pkg = sys.modules.get("util") or types.ModuleType("util")
        util_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "util")]
        misc = types.ModuleType("util.misc")

        def inverse_sigmoid(x, eps: float = 1e-3):
            x = x.clamp(min=0, max=1)
            return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))

        misc.inverse_sigmoid = inverse_sigmoid
        sys.modules["util"] = util_pkg
        sys.modules["util.misc"] = misc
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.bricks.ms_deform_attn import (MultiScaleDeformableAttention,
                                                  multi_scale_deformable_attn_pytorch)
        from models.bricks.relation_transformer import PositionRelationEmbedding, box_rel_encoding
    return multi_scale_deformable_attn_pytorch, PositionRelationEmbedding, box_rel_encoding, MultiScaleDeformableAttention
