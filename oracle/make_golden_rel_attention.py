"""Generate tests/golden/relattn_*.npz from the REFERENCE (build container only) -- fixtures for row N1.

TEST INFRASTRUCTURE ONLY.  Run as ``python -m oracle.make_golden_rel_attention``.  For seeded inputs it evaluates,
in float64 on the CPU, exactly what the reference's decoder does around its self-attention:

    pos_relation = PositionRelationEmbedding(16, 8)(src_boxes, tgt_boxes).flatten(0, 1)     relation_transformer.py:372
    pos_relation.masked_fill_(attn_mask, -inf)                                              :373-374
    out = nn.MultiheadAttention(256, 8, batch_first=True)(q_pos, q_pos, query, attn_mask=pos_relation)[0]   :447-455

and stores inputs, the module parameters, the output and the autograd gradients of ``(out * g).sum()``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from relation_detr_b200 import workloads as wl  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
# name, B, N, denoising (max_gt, groups) or None, seed
CASES = [("relattn_plain", 2, 23, None, 0), ("relattn_cdn", 1, 40, (3, 4), 1)]


def main():
    _, PositionRelationEmbedding, _, _ = ref_import.load()
    E, H = 256, 8
    # one set of module parameters for all cases, stored once; values are float32-representable so that the
    # file stays small (the modules themselves run in float64)
    torch.manual_seed(0)
    rel = PositionRelationEmbedding(16, H)
    mha = torch.nn.MultiheadAttention(E, H, dropout=0.0, batch_first=True)
    torch.nn.init.xavier_uniform_(mha.in_proj_weight)
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.02)
        mha.out_proj.bias.normal_(0, 0.02)
    np.savez_compressed(
        os.path.join(GOLDEN, "relattn_weights.npz"), in_proj_weight=mha.in_proj_weight.detach().numpy(),
        in_proj_bias=mha.in_proj_bias.detach().numpy(), out_proj_weight=mha.out_proj.weight.detach().numpy(),
        out_proj_bias=mha.out_proj.bias.detach().numpy(), rel_weight=rel.pos_proj[0].weight.detach().numpy().reshape(H, 64),
        rel_bias=rel.pos_proj[0].bias.detach().numpy())
    rel, mha = rel.double(), mha.double()
    for name, B, N, dn, seed in CASES:
        g = torch.Generator().manual_seed(seed)
        rel.zero_grad()
        mha.zero_grad()
        query = torch.randn(B, N, E, generator=g, dtype=torch.float64, requires_grad=True)
        query_pos = torch.randn(B, N, E, generator=g, dtype=torch.float64)
        src = wl.make_boxes(B, N, seed + 10, "cpu").double()
        tgt = wl.make_boxes(B, N, seed + 20, "cpu").double()
        mask = None
        if dn is not None:
            mask = wl.cdn_attn_mask(N - dn[0] * dn[1], dn[0], dn[1], "cpu")   # denoising rows first (denoising.py:66-78)
        qp = query + query_pos
        pos_relation = rel(src, tgt).flatten(0, 1)
        if mask is not None:
            assert mask.shape == (N, N), mask.shape
            pos_relation.masked_fill_(mask, float("-inf"))
        out = mha(query=qp, key=qp, value=query, attn_mask=pos_relation, need_weights=False)[0]
        gout = torch.randn(out.shape, generator=g, dtype=torch.float64)
        (out * gout).sum().backward()
        np.savez_compressed(
            os.path.join(GOLDEN, name + ".npz"), query=query.detach().numpy(), query_pos=query_pos.numpy(), src_boxes=src.numpy(),
            tgt_boxes=tgt.numpy(), attn_mask=(mask.numpy() if mask is not None else np.zeros((0, 0), bool)),
            out=out.detach().numpy(), grad_out=gout.numpy(), grad_query=query.grad.numpy(),
            grad_in_proj_bias=mha.in_proj_bias.grad.numpy(), grad_rel_weight=rel.pos_proj[0].weight.grad.numpy().reshape(H, 64),
            grad_rel_bias=rel.pos_proj[0].bias.grad.numpy())
        print(name, tuple(out.shape), "mask" if mask is not None else "no mask")


if __name__ == "__main__":
    main()
