"""Swap the B200 operators into an importable copy of the reference repository.

``install()`` rebinds ``MultiScaleDeformableAttention`` / ``MultiScaleDeformableAttnFunction`` /
``PositionRelationEmbedding`` inside the reference's own modules, so that
``models/detectors/relation_detr`` and the DINO++ / Deformable-DETR++ configs build with the new
path without editing them.  Call it before the config file is executed (configs instantiate the
model at import time, ``util/lazy_load.py``).  See INTEGRATION.md.
"""
from __future__ import annotations

import importlib
import sys

from . import matcher, modules, ops

# modules of the reference that bind the two classes by name at import time
_MSDA_USERS = (
    "models.bricks.ms_deform_attn",
    "models.bricks.relation_transformer",
    "models.bricks.dino_transformer",
    "models.bricks.deformable_transformer",
    "models.bricks.dab_transformer",
    "models.bricks.dn_transformer",
)


# modules of the reference that bind HungarianMatcher by name (the configs import it from the first one)
_MATCHER_USERS = ("models.matcher.hungarian_matcher",)


def install(reference_root: str | None = None, strict: bool = False, matcher_too: bool = True) -> list:
    """Returns the list of ``module.attribute`` names that were rebound.  ``matcher_too`` also replaces
    ``HungarianMatcher`` (device-resident matching): its index tensors are CUDA
    tensors, which every use in ``models/bricks/set_criterion.py`` accepts."""
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    rebound = []
    for name in _MSDA_USERS:
        try:
            mod = importlib.import_module(name)
        except Exception:
            if strict:
                raise
            continue
        for attr, repl in (("MultiScaleDeformableAttention", modules.MultiScaleDeformableAttention),
                           ("MultiScaleDeformableAttnFunction", ops.MultiScaleDeformableAttnFunction),
                           ("PositionRelationEmbedding", modules.PositionRelationEmbedding)):
            if hasattr(mod, attr):
                setattr(mod, attr, repl)
                rebound.append(f"{name}.{attr}")
    for name in _MATCHER_USERS if matcher_too else ():
        try:
            mod = importlib.import_module(name)
        except Exception:
            if strict:
                raise
            continue
        mod.HungarianMatcher = matcher.HungarianMatcher
        rebound.append(f"{name}.HungarianMatcher")
    return rebound
