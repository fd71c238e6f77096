"""Generate tests/golden/twostage_*.npz from the REFERENCE (build container only) -- fixtures for row N4's selection.

TEST INFRASTRUCTURE ONLY.  Run as ``python -m oracle.make_golden_two_stage``.  It runs the reference's own
``RelationTransformer.forward`` (``models/bricks/relation_transformer.py:59-151``, training mode, so both the main and
the hybrid selection execute) around a pass-through encoder and an inert decoder, records what went into the two
selections (the class heads' outputs and ``bbox_head(output_memory) + output_proposals``) with forward hooks, and stores
them next to what ``forward`` returned (``enc_outputs_class``, ``enc_outputs_coord``, ``hybrid_enc_class``,
``hybrid_enc_coord``).  One image is padded, so part of the proposals is +inf and many score ties exist below the cut.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
# name, B, level shapes, classes, main k, hybrid k, seed
CASES = [("twostage_small", 2, ((9, 13), (5, 7)), 7, 11, 19, 0), ("twostage_pad", 3, ((12, 20), (6, 10), (3, 5)), 13, 40, 25, 1)]


class _PassThroughEncoder(nn.Module):
    embed_dim = 32

    def forward(self, query, **kwargs):
        return query


class _InertDecoder(nn.Module):
    def forward(self, **kwargs):
        return None, None


def main():
    ref_import.load()
    from models.bricks.relation_transformer import RelationTransformer

    for name, B, levels, C, k_main, k_hyb, seed in CASES:
        torch.manual_seed(seed)
        model = RelationTransformer(_PassThroughEncoder(), _InertDecoder(), num_classes=C, num_feature_levels=len(levels),
                                    two_stage_num_proposals=k_main, hybrid_num_proposals=k_hyb).train()
        with torch.no_grad():   # the reference zero-initialises the box heads' last layer: make them matter
            for head in (model.encoder_bbox_head, model.hybrid_bbox_head):
                head.layers[-1].weight.normal_(0, 0.5)
                head.layers[-1].bias.normal_(0, 0.5)
        rec = {}
        model.encoder_class_head.register_forward_hook(lambda m, i, o: rec.__setitem__("main_class", o.detach().clone()))
        model.hybrid_class_head.register_forward_hook(lambda m, i, o: rec.__setitem__("hybrid_class", o.detach().clone()))
        model.encoder_bbox_head.register_forward_hook(lambda m, i, o: rec.__setitem__("main_box", o.detach().clone()))
        model.hybrid_bbox_head.register_forward_hook(lambda m, i, o: rec.__setitem__("hybrid_box", o.detach().clone()))
        inner = model.get_encoder_output

        def recording(memory, proposals, mask, inner=inner, rec=rec):
            out = inner(memory, proposals, mask)
            rec["proposals"] = out[1].detach().clone()
            return out

        model.get_encoder_output = recording
        g = torch.Generator().manual_seed(seed)
        feats = [torch.randn(B, 32, h, w, generator=g) for h, w in levels]
        pos = [torch.randn(B, 32, h, w, generator=g) for h, w in levels]
        masks = [torch.zeros(B, h, w, dtype=torch.bool) for h, w in levels]
        for m in masks:   # the last image is padded on the right and at the bottom (a quarter of each side)
            m[-1, :, -max(1, m.shape[2] // 4):] = True
            m[-1, -max(1, m.shape[1] // 4):, :] = True
        with torch.no_grad():
            out = model(feats, masks, pos)
        _, _, enc_class, enc_coord, _, _, hyb_class, hyb_coord = out
        np.savez_compressed(
            os.path.join(GOLDEN, name + ".npz"),
            main_class=rec["main_class"].numpy(), main_coord_unact=(rec["main_box"] + rec["proposals"]).numpy(),
            hybrid_class=rec["hybrid_class"].numpy(), hybrid_coord_unact=(rec["hybrid_box"] + rec["proposals"]).numpy(),
            k_main=np.int64(k_main), k_hybrid=np.int64(k_hyb),
            expected_main_class=enc_class.numpy(), expected_main_coord=enc_coord.numpy(),
            expected_hybrid_class=hyb_class.numpy(), expected_hybrid_coord=hyb_coord.numpy())
        print(name, {k: tuple(v.shape) for k, v in rec.items()}, "->", tuple(enc_class.shape), tuple(hyb_class.shape),
              "invalid rows:", int(torch.isinf(rec["proposals"]).any(-1).sum()))


if __name__ == "__main__":
    main()
