"""Builds the reference's own models from its own classes (``baseline/_ref`` or ``/root/reference``).

``build_relation_detr_r50`` restates the constructor calls of
``configs/relation_detr/relation_detr_resnet50_800_1333.py`` (upstream) with ``weights=False`` for the backbone:
executing the config file itself downloads ImageNet weights (``models/backbones/resnet.py:428-432``), and there is
no network here.  Everything else -- classes, arguments, loss weights -- is the config's.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root() -> str | None:
    for cand in (os.path.join(HERE, "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "models", "bricks")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def activate() -> str:
    """Puts the reference tree on sys.path (after registering the stand-ins for its missing imports) and imports
    ``models.bricks.ms_deform_attn`` once with its import-time JIT build short-circuited.

    Upstream builds its CUDA extension at import time whenever a GPU is visible (``ms_deform_attn.py:14-26``); against
    this image's torch 2.11 that build runs nvcc for minutes and then fails, after which upstream
    silently uses ``multi_scale_deformable_attn_pytorch``.  We make the doomed build fail immediately instead --
    ``torch.utils.cpp_extension.load`` raises for the duration of that one import -- which leaves the reference in
    exactly the state it reaches on its own (``_C is None``).  No reference file is edited.
    ``set_reference_extension("prebuilt")`` can then hand it its own kernel, compiled from its unmodified sources by
    ``oracle/build_ref_cuda.py``."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found: run `python baseline/install_reference.py` in the build container")
    from . import stubs

    stubs.install()
    if root not in sys.path:
        sys.path.insert(0, root)
    if "models.bricks.ms_deform_attn" not in sys.modules:
        import warnings

        import torch.utils.cpp_extension as cpp_ext

        real_load = cpp_ext.load

        def _no_jit(*a, **k):
            raise RuntimeError("JIT build skipped by baseline/refmodel.py: upstream's extension does not compile against torch 2.11 "
                               "(SURVEY.md F1); the reference falls back to multi_scale_deformable_attn_pytorch as it does on its own")

        cpp_ext.load = _no_jit
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                import models.bricks.ms_deform_attn  # noqa: F401
        finally:
            cpp_ext.load = real_load
    return root


def set_reference_extension(mode: str = "none") -> str:
    """Which MSDA path the UNMODIFIED reference modules take on the GPU: ``"none"`` = its grid_sample path (what
    upstream effectively runs on this image), ``"prebuilt"`` = its own CUDA kernel from ``oracle/_ref`` (unmodified
    sources recompiled for sm_100a).  Sets the module global the reference itself consults (``ms_deform_attn.py:358``);
    returns the mode actually in effect."""
    activate()
    import models.bricks.ms_deform_attn as ref_msda

    if mode == "prebuilt":
        root = os.path.dirname(HERE)
        if root not in sys.path:
            sys.path.insert(0, root)
        from oracle import build_ref_cuda

        ext = build_ref_cuda.load_prebuilt()
        ref_msda._C = ext
        return "prebuilt" if ext is not None else "none"
    ref_msda._C = None
    return "none"


def build_relation_detr_r50(num_classes: int = 91, num_queries: int = 900, hybrid_num_proposals: int = 1500,
                            enc_layers: int = 6, dec_layers: int = 6, num_feature_levels: int = 4, denoising_nums: int = 100,
                            min_size: int = 800, max_size: int = 1333, backbone_name: str = "resnet50"):
    """RelationDETR R50 as configs/relation_detr/relation_detr_resnet50_800_1333.py builds it, random init.
    Call ``relation_detr_b200.install.install()`` BEFORE this to get the B200 operators inside.
    ``backbone_name="focalnet_large_lrf_fl4"`` swaps in the backbone line of
    configs/relation_detr/relation_detr_focalnet_large_lrf_fl4_1200_2000.py:35 (see ``build_relation_detr_focal_l``)."""
    activate()
    from torch import nn

    from models.backbones.resnet import ResNetBackbone
    from models.bricks.misc import FrozenBatchNorm2d
    from models.bricks.position_encoding import PositionEmbeddingSine
    from models.bricks.post_process import PostProcess
    from models.bricks import relation_transformer as rt
    from models.bricks.set_criterion import HybridSetCriterion
    from models.detectors.relation_detr import RelationDETR
    from models.matcher import hungarian_matcher as hm
    from models.necks.channel_mapper import ChannelMapper

    embed_dim, num_heads, dim_feedforward, hybrid_assign = 256, 8, 2048, 6
    position_embedding = PositionEmbeddingSine(embed_dim // 2, temperature=10000, normalize=True, offset=-0.5)
    if backbone_name == "resnet50":
        backbone = ResNetBackbone("resnet50", weights=False, norm_layer=FrozenBatchNorm2d, return_indices=(1, 2, 3), freeze_indices=(0,))
    else:
        from models.backbones.focalnet import FocalNetBackbone

        backbone = FocalNetBackbone(backbone_name, weights=False, return_indices=(0, 1, 2, 3))
    neck = ChannelMapper(in_channels=backbone.num_channels, out_channels=embed_dim, num_outs=num_feature_levels)
    transformer = rt.RelationTransformer(
        encoder=rt.RelationTransformerEncoder(
            encoder_layer=rt.RelationTransformerEncoderLayer(
                embed_dim=embed_dim, n_heads=num_heads, dropout=0.0, activation=nn.ReLU(inplace=True),
                n_levels=num_feature_levels, n_points=4, d_ffn=dim_feedforward),
            num_layers=enc_layers),
        decoder=rt.RelationTransformerDecoder(
            decoder_layer=rt.RelationTransformerDecoderLayer(
                embed_dim=embed_dim, n_heads=num_heads, dropout=0.0, activation=nn.ReLU(inplace=True),
                n_levels=num_feature_levels, n_points=4, d_ffn=dim_feedforward),
            num_layers=dec_layers, num_classes=num_classes),
        num_classes=num_classes, num_feature_levels=num_feature_levels, two_stage_num_proposals=num_queries,
        hybrid_num_proposals=hybrid_num_proposals)
    matcher = hm.HungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, focal_alpha=0.25, focal_gamma=2.0)
    weight_dict = {"loss_class": 1, "loss_bbox": 5, "loss_giou": 2}
    weight_dict.update({"loss_class_dn": 1, "loss_bbox_dn": 5, "loss_giou_dn": 2})
    aux = {}
    for i in range(transformer.decoder.num_layers - 1):
        aux.update({k + f"_{i}": v for k, v in weight_dict.items()})
    weight_dict.update(aux)
    weight_dict.update({"loss_class_enc": 1, "loss_bbox_enc": 5, "loss_giou_enc": 2})
    weight_dict.update({k + "_hybrid": v for k, v in weight_dict.items()})
    criterion = HybridSetCriterion(num_classes=num_classes, matcher=matcher, weight_dict=weight_dict, alpha=0.25, gamma=2.0)
    postprocessor = PostProcess(select_box_nums_for_evaluation=300)
    model = RelationDETR(backbone=backbone, neck=neck, position_embedding=position_embedding, transformer=transformer,
                         criterion=criterion, postprocessor=postprocessor, num_classes=num_classes, num_queries=num_queries,
                         hybrid_assign=hybrid_assign, denoising_nums=denoising_nums, min_size=min_size, max_size=max_size)
    return model, weight_dict


def build_relation_detr_focal_l(enc_layers: int = 6, dec_layers: int = 6):
    """BASELINE configs[4]: RelationDETR with the FocalNet-L backbone, 5 feature levels, 1200x2000 images and
    denoising_nums = 1000, as configs/relation_detr/relation_detr_focalnet_large_lrf_fl4_1200_2000.py builds it
    (:24 num_feature_levels = 5, :35 backbone with weights=False, :101-103 denoising_nums / min_size / max_size)."""
    return build_relation_detr_r50(enc_layers=enc_layers, dec_layers=dec_layers, num_feature_levels=5, denoising_nums=1000,
                                   min_size=1200, max_size=2000, backbone_name="focalnet_large_lrf_fl4")


def synthetic_batch(batch: int, device, seed: int = 0, height: int = 800, width: int = 1333, boxes_per_image: int = 10,
                    num_classes: int = 91):
    """Fixed-size synthetic images and targets: ``randn(3,H,W)`` images, ``boxes_per_image``
    random valid xyxy boxes in pixels and labels in [0, num_classes)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    images, targets = [], []
    for _ in range(batch):
        images.append(torch.randn((3, height, width), generator=g).to(device))
        cx = torch.rand((boxes_per_image,), generator=g) * (width * 0.8) + width * 0.1
        cy = torch.rand((boxes_per_image,), generator=g) * (height * 0.8) + height * 0.1
        w = torch.rand((boxes_per_image,), generator=g) * (width * 0.15) + 8
        h = torch.rand((boxes_per_image,), generator=g) * (height * 0.15) + 8
        boxes = torch.stack([(cx - w / 2).clamp(min=0), (cy - h / 2).clamp(min=0), (cx + w / 2).clamp(max=width), (cy + h / 2).clamp(max=height)], -1)
        labels = torch.randint(0, num_classes, (boxes_per_image,), generator=g)
        targets.append({"boxes": boxes.to(device), "labels": labels.to(device)})
    return images, targets
