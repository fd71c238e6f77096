"""Swap the B200 operators into an importable copy of the reference repository.

``install()`` rebinds ``MultiScaleDeformableAttention`` / ``MultiScaleDeformableAttnFunction`` /
``PositionRelationEmbedding`` inside the reference's own modules, so that
``models/detectors/relation_detr`` and the DINO++ / Deformable-DETR++ configs build with the new
path without editing them.  Call it before the config file is executed (configs instantiate the
model at import time, ``util/lazy_load.py``).  See INTEGRATION.md.

It fails loudly: a module of the reference that cannot be imported raises (the reference itself falls back
silently when its extension does not build, ``ms_deform_attn.py:23-26`` -- exactly the failure mode this
package refuses to have).  ``strict=False`` restores the lenient behaviour and returns what it skipped.
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

_saved = {}  # "module.attr" -> the reference's own object, for uninstall()


class InstallReport(list):
    """The rebound ``module.attribute`` names (a list, as before) plus ``skipped``: modules that failed to import
    (only ever non-empty with ``strict=False``)."""

    def __init__(self, rebound=(), skipped=()):
        super().__init__(rebound)
        self.skipped = list(skipped)


class _NNProxy:
    """Stands in for the name ``nn`` inside ``models.bricks.relation_transformer`` (``from torch import nn``): every
    attribute is ``torch.nn``'s except ``MultiheadAttention``, so that ``RelationTransformerDecoderLayer.__init__``
    (``relation_transformer.py:401``) builds the fused self-attention without the file being edited."""

    def __getattr__(self, name):
        import torch.nn

        if name == "MultiheadAttention":
            return modules.RelationMultiheadAttention
        return getattr(torch.nn, name)


class _TorchProxy:
    """Stands in for the name ``torch`` inside ``models.bricks.relation_transformer``: every attribute is ``torch``'s except
    ``topk``, so that the two-stage selection of ``RelationTransformer.forward`` (``relation_transformer.py:94, :109``:
    ``torch.topk(enc_outputs_class.max(-1)[0], topk, dim=1)``) runs as one CTA per image (``rdetr_topk_rows``) without the
    file being edited.  Calls the kernel does not cover (other ranks, dims, dtypes, k > 4096, CPU tensors) are torch's own."""

    def __getattr__(self, name):
        import torch

        return getattr(torch, name)

    @staticmethod
    def topk(input, k, dim=-1, largest=True, sorted=True, **kwargs):
        import torch

        if (isinstance(input, torch.Tensor) and input.is_cuda and input.dim() == 2 and dim in (1, -1) and largest and sorted
                and not kwargs and input.dtype in (torch.float32, torch.bfloat16, torch.float16) and 0 < int(k) <= min(4096, input.shape[1])):
            values, indices = ops.topk_rows(input.float(), int(k))   # widening is exact: same order
            return torch.return_types.topk((input.gather(1, indices) if input.dtype != torch.float32 else values, indices))
        return torch.topk(input, k, dim=dim, largest=largest, sorted=sorted, **kwargs)


def install(reference_root: str | None = None, strict: bool = True, matcher_too: bool = True,
            fused_attention: bool = False, fused_memory: bool = False, fused_topk: bool = False) -> InstallReport:
    """Returns the ``module.attribute`` names that were rebound.  ``matcher_too`` also replaces
    ``HungarianMatcher`` (device-resident matching): its index tensors are CUDA
    tensors, which every use in ``models/bricks/set_criterion.py`` accepts.  ``fused_attention`` (row N1) makes the
    relation decoder's self-attention generate the position-relation bias inside the attention kernel:
    ``PositionRelationEmbedding`` hands out a lazy handle and the decoder layer's ``nn.MultiheadAttention`` becomes
    ``RelationMultiheadAttention`` (same parameters, same state-dict keys).  ``fused_memory`` (row N4) replaces
    ``RelationTransformerEncoder`` by a subclass whose ``memory_fusion`` input Linear reads the encoder states in place
    (tcgen05 GEMM) instead of concatenating them.  ``fused_topk`` (row N4, second half) routes the two ``torch.topk`` calls of
    ``RelationTransformer.forward`` to ``rdetr_topk_rows`` (same indices; equal scores in ascending index order)."""
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    report = InstallReport()

    def rebind(mod, name, attr, repl):
        key = f"{name}.{attr}"
        if key not in _saved:
            _saved[key] = getattr(mod, attr)
        setattr(mod, attr, repl)
        report.append(key)

    # Import every user BEFORE rebinding anything: `relation_transformer` does `from models.bricks.ms_deform_attn import
    # MultiScaleDeformableAttention` at import time, so importing it after `ms_deform_attn` has been rebound would make
    # this package's class look like the reference's own one -- uninstall() would then "restore" ours (found by the
    # leak guard in tests/conftest.py when install() happened to be the first importer of the reference).
    users = []
    for name in _MSDA_USERS:
        try:
            users.append((name, importlib.import_module(name)))
        except Exception as e:
            if strict:
                raise RuntimeError(f"relation_detr_b200.install: cannot import the reference module {name!r} "
                                   f"({type(e).__name__}: {e}); nothing of it was rebound") from e
            report.skipped.append(f"{name} ({type(e).__name__}: {e})")
    for name, mod in users:
        for attr, repl in (("MultiScaleDeformableAttention", modules.MultiScaleDeformableAttention),
                           ("MultiScaleDeformableAttnFunction", ops.MultiScaleDeformableAttnFunction),
                           ("PositionRelationEmbedding", modules.PositionRelationEmbedding)):
            if hasattr(mod, attr):
                rebind(mod, name, attr, repl)
    for name in _MATCHER_USERS if matcher_too else ():
        try:
            mod = importlib.import_module(name)
        except Exception as e:
            if strict:
                raise RuntimeError(f"relation_detr_b200.install: cannot import the reference module {name!r} "
                                   f"({type(e).__name__}: {e})") from e
            report.skipped.append(f"{name} ({type(e).__name__}: {e})")
            continue
        rebind(mod, name, "HungarianMatcher", matcher.HungarianMatcher)
    if fused_attention:
        try:
            mod = importlib.import_module("models.bricks.relation_transformer")
        except Exception as e:
            raise RuntimeError(f"relation_detr_b200.install: cannot import models.bricks.relation_transformer ({type(e).__name__}: {e})") from e
        rebind(mod, "models.bricks.relation_transformer", "nn", _NNProxy())
        if "PositionRelationEmbedding.lazy" not in _saved:
            _saved["PositionRelationEmbedding.lazy"] = modules.PositionRelationEmbedding.lazy
        modules.PositionRelationEmbedding.lazy = True
    if fused_memory:
        try:
            mod = importlib.import_module("models.bricks.relation_transformer")
        except Exception as e:
            raise RuntimeError(f"relation_detr_b200.install: cannot import models.bricks.relation_transformer ({type(e).__name__}: {e})") from e
        base = _saved.get("models.bricks.relation_transformer.RelationTransformerEncoder", mod.RelationTransformerEncoder)
        rebind(mod, "models.bricks.relation_transformer", "RelationTransformerEncoder", modules.make_fused_encoder(base))
    if fused_topk:
        try:
            mod = importlib.import_module("models.bricks.relation_transformer")
        except Exception as e:
            raise RuntimeError(f"relation_detr_b200.install: cannot import models.bricks.relation_transformer ({type(e).__name__}: {e})") from e
        rebind(mod, "models.bricks.relation_transformer", "torch", _TorchProxy())
    if strict and not report:
        raise RuntimeError("relation_detr_b200.install: no name of the reference was rebound")
    return report


def uninstall() -> list:
    """Puts the reference's own classes back (models built in between keep whatever they were built with)."""
    restored = []
    if "PositionRelationEmbedding.lazy" in _saved:
        modules.PositionRelationEmbedding.lazy = _saved.pop("PositionRelationEmbedding.lazy")
    for key, obj in list(_saved.items()):
        name, attr = key.rsplit(".", 1)
        mod = sys.modules.get(name)
        if mod is not None:
            setattr(mod, attr, obj)
            restored.append(key)
        del _saved[key]
    return restored
