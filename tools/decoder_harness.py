"""Measurement / integration harness for BASELINE.json configs[2]: a relation decoder stack that CALLS
the two hot-path operators the way the reference's decoder does.

This is not product code and not a copy of the reference: it is the smallest caller that reproduces the
call pattern of ``RelationTransformerDecoder`` / ``RelationTransformerDecoderLayer``
(upstream models/bricks/relation_transformer.py:279-478) -- per layer: self-attention whose additive
mask is the position-relation bias, MSDA cross-attention on 4-d reference boxes, FFN; between layers:
box refinement and ``REL(previous boxes, new boxes)`` with the denoising mask applied in place.  The
reference tree does not exist on the GPU box, hence the restatement.  ``impl="ours"`` uses the drop-in
modules of ``relation_detr_b200``; ``impl="oracle"`` routes the same weights through the oracle
(grid_sample MSDA + eager relation embedding), i.e. the path the unmodified reference runs on this image.
"""
from __future__ import annotations

import copy
import math
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relation_detr_b200 as rd  # noqa: E402


def mlp(inp, hidden, out, layers):
    dims = [inp] + [hidden] * (layers - 1) + [out]
    mods = []
    for i in range(layers):
        mods.append(nn.Linear(dims[i], dims[i + 1]))
        if i < layers - 1:
            mods.append(nn.ReLU(inplace=True))
    return nn.Sequential(*mods)


def sine_embed(pos, feats=128, temperature=10000.0, scale=2 * math.pi):
    """Sine embedding of [..., 4] boxes with x/y exchanged (position_encoding.py:115-138 semantics)."""
    dim_t = temperature ** (torch.arange(feats // 2, dtype=torch.float32, device=pos.device) * 2 / feats)
    ang = pos.unsqueeze(-1) * scale / dim_t
    emb = torch.stack((ang.sin(), ang.cos()), dim=-1).flatten(-2)  # [..., 4, feats]
    # (y, x, w, h) order; sliced rather than index_select so the step stays CUDA-graph capturable
    return torch.cat((emb[..., 1:2, :], emb[..., 0:1, :], emb[..., 2:, :]), dim=-2).flatten(-2)


def inverse_sigmoid(x, eps=1e-3):
    x = x.clamp(0, 1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


class OracleMSDA(rd.MultiScaleDeformableAttention):
    """Same parameters and prologue as the drop-in module, sampling done by the oracle (grid_sample)."""

    def forward(self, query, reference_points, value, spatial_shapes, level_start_index, key_padding_mask):
        from oracle import torch_port

        B, Nq, _ = query.shape
        S = value.shape[1]
        M, L, P = self.num_heads, self.num_levels, self.num_points
        value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(B, S, M, self.embed_dim // M)
        off = self.sampling_offsets(query).view(B, Nq, M, L, P, 2)
        w = self.attention_weights(query).view(B, Nq, M, L * P).softmax(-1).view(B, Nq, M, L, P)
        if reference_points.shape[-1] == 2:
            wh = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
            loc = reference_points[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
        else:
            loc = reference_points[:, :, None, :, None, :2] + off / P * reference_points[:, :, None, :, None, 2:] * 0.5
        return self.output_proj(torch_port.msda_grid_sample(value, spatial_shapes, loc, w))


class OracleREL(rd.PositionRelationEmbedding):
    def forward(self, src_boxes, tgt_boxes=None, attn_mask=None):
        from oracle import torch_port

        out = torch_port.rel_eager(src_boxes, tgt_boxes, self.pos_proj[0].weight, self.pos_proj[0].bias)
        return out.clone() if attn_mask is None else out.masked_fill(attn_mask, float("-inf"))


class DecoderLayer(nn.Module):
    def __init__(self, impl, d=256, ffn=1024, heads=8, levels=4, points=4):
        super().__init__()
        msda = rd.MultiScaleDeformableAttention if impl == "ours" else OracleMSDA
        self.cross_attn = msda(d, levels, heads, points)
        self.self_attn = nn.MultiheadAttention(d, heads, dropout=0.0, batch_first=True)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)
        self.linear1, self.linear2 = nn.Linear(d, ffn), nn.Linear(ffn, d)

    def forward(self, query, query_pos, reference_points, value, spatial_shapes, level_start_index, self_attn_mask):
        qk = query + query_pos
        query = self.norm2(query + self.self_attn(qk, qk, query, attn_mask=self_attn_mask, need_weights=False)[0])
        query = self.norm1(query + self.cross_attn(query + query_pos, reference_points, value, spatial_shapes,
                                                  level_start_index, None))
        return self.norm3(query + self.linear2(torch.relu(self.linear1(query))))


class RelationDecoder(nn.Module):
    def __init__(self, impl="ours", layers=6, d=256, heads=8, levels=4, classes=91):
        super().__init__()
        self.layers = nn.ModuleList([DecoderLayer(impl, d, 1024, heads, levels) for _ in range(layers)])
        self.ref_point_head = mlp(2 * d, d, d, 2)
        self.query_scale = mlp(d, d, d, 2)
        self.class_head = nn.ModuleList([nn.Linear(d, classes) for _ in range(layers)])
        self.bbox_head = nn.ModuleList([mlp(d, d, 4, 3) for _ in range(layers)])
        self.norm = nn.LayerNorm(d)
        rel = rd.PositionRelationEmbedding if impl == "ours" else OracleREL
        self.position_relation_embedding = rel(16, heads)
        self.d = d

    def forward(self, query, reference_points, value, spatial_shapes, level_start_index, valid_ratios, attn_mask=None,
                skip_relation=False):
        classes, coords = [], []
        scale = torch.cat([valid_ratios, valid_ratios], -1)[:, None]
        pos_relation = attn_mask
        tgt_boxes = None
        for i, layer in enumerate(self.layers):
            ref_in = reference_points.detach()[:, :, None] * scale
            query_pos = self.ref_point_head(sine_embed(ref_in[:, :, 0, :], self.d // 2))
            if i:
                query_pos = query_pos * self.query_scale(query)
            query = layer(query, query_pos, ref_in, value, spatial_shapes, level_start_index, pos_relation)
            normed = self.norm(query)
            coord = (self.bbox_head[i](normed) + inverse_sigmoid(reference_points)).sigmoid()
            classes.append(self.class_head[i](normed))
            coords.append(coord)
            if i == len(self.layers) - 1:
                break
            if not skip_relation:
                src_boxes = tgt_boxes if i >= 1 else reference_points
                tgt_boxes = coord
                pos_relation = self.position_relation_embedding(src_boxes, tgt_boxes).flatten(0, 1)
                if attn_mask is not None:
                    pos_relation.masked_fill_(attn_mask, float("-inf"))
            reference_points = (self.bbox_head[i](query) + inverse_sigmoid(reference_points.detach())).sigmoid()
        return torch.stack(classes), torch.stack(coords)


def build_pair(seed=0, device="cuda", **kw):
    """(ours, oracle) decoders with identical weights."""
    torch.manual_seed(seed)
    ours = RelationDecoder("ours", **kw).to(device)
    with torch.no_grad():  # non-trivial sampling offsets / attention logits / box heads
        for layer in ours.layers:
            layer.cross_attn.sampling_offsets.weight.normal_(0, 0.01)
            layer.cross_attn.attention_weights.weight.normal_(0, 0.02)
        for head in ours.bbox_head:
            head[-1].weight.mul_(0.1)
    oracle = RelationDecoder("oracle", **kw).to(device)
    oracle.load_state_dict(copy.deepcopy(ours.state_dict()), strict=True)
    return ours, oracle


def make_inputs(batch, num_queries, dn_rows, levels, seed=0, device="cuda"):
    from relation_detr_b200 import workloads

    g = torch.Generator(device=device).manual_seed(seed)
    ss, lsi = workloads.shape_tensors(levels, device)
    S = int(ss.prod(1).sum())
    n = num_queries + dn_rows
    query = torch.randn((batch, n, 256), device=device, generator=g)
    refs = workloads.make_boxes(batch, n, seed + 5, device)
    refs[..., 2:] = refs[..., 2:] * 0.6 + 0.05
    memory = torch.randn((batch, S, 256), device=device, generator=g)
    valid = torch.ones((batch, len(levels), 2), device=device)
    mask = workloads.cdn_attn_mask(num_queries, dn_rows // 20, 20, device) if dn_rows else None
    return dict(query=query, reference_points=refs, value=memory, spatial_shapes=ss, level_start_index=lsi,
                valid_ratios=valid, attn_mask=mask)


def time_block(model, inp, hybrid_inp=None, warmup=2, iters=5):
    """fwd+bwd of sum(outputs) for the main pass (with REL) and, if given, the hybrid pass (skip_relation)."""
    def step():
        model.zero_grad(set_to_none=True)
        c, b = model(**inp)
        loss = c.sum() + b.sum()
        if hybrid_inp is not None:
            c2, b2 = model(**hybrid_inp, skip_relation=True)
            loss = loss + c2.sum() + b2.sum()
        loss.backward()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    import json

    from relation_detr_b200 import workloads

    ours, oracle = build_pair(0)
    main_inp = make_inputs(8, 900, 200, workloads.LEVELS_800_1333, 0)
    hyb = make_inputs(8, 1500, 0, workloads.LEVELS_800_1333, 1)
    res = {"config": "configs[2]: 6 layers, B=8, main pass N=900+200 dn rows with REL + CDN mask, hybrid pass N=1500 without REL, S=22323, fwd+bwd of sum(outputs), fp32"}
    res["ours_ms"] = time_block(ours, main_inp, hyb)
    res["reference_path_ms"] = time_block(oracle, main_inp, hyb)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        res["ours_bf16_autocast_ms"] = time_block(ours, main_inp, hyb)
    ours.position_relation_embedding.fast_math = False
    res["ours_rel_exact_ms"] = time_block(ours, main_inp, hyb)
    res["speedup_vs_reference_path"] = res["reference_path_ms"] / res["ours_ms"]
    print(json.dumps(res, indent=1))
