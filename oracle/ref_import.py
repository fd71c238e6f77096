"""Import the two reference operators from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing that runs
there may call this; it is used by ``make_golden.py`` and by CPU tests that skip when the reference
tree is absent.  The third-party packages the reference imports but this image lacks are replaced by the inert
stand-ins of ``baseline/stubs.py``; the reference's own modules are imported unmodified.
"""
from __future__ import annotations

import os
import sys
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
    # the reference's util.misc drags in packages that are not installed here: register the
    # inert stand-ins of baseline/stubs.py for them and import the REAL util.misc (one mechanism for every
    # consumer of the reference tree in this repository: golden generation, CPU tests, the reference arm)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from baseline import stubs

    stubs.install()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.bricks.ms_deform_attn import (MultiScaleDeformableAttention,
                                                  multi_scale_deformable_attn_pytorch)
        from models.bricks.relation_transformer import PositionRelationEmbedding, box_rel_encoding
    return multi_scale_deformable_attn_pytorch, PositionRelationEmbedding, box_rel_encoding, MultiScaleDeformableAttention
