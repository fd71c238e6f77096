"""Torch restatement of the reference's CPU path for the two operators.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  This is the path the unmodified reference
runs whenever its CUDA extension is unavailable (``models/bricks/ms_deform_attn.py:367-370``), i.e.
per-level ``F.grid_sample`` + weighted sum, and the eager sin/cos + 1x1 conv relation embedding
(``models/bricks/relation_transformer.py:481-532``).  It issues the same ATen calls in the same
order as the reference, so timing it on the host cores is timing the reference's CPU path
(``cpu_baseline.kind == "port"``).  It is differentiable, so autograd supplies the backward exactly
as it does for the reference.  Pinned against ``tests/golden/`` (see ``make_golden.py``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def msda_grid_sample(value, spatial_shapes, sampling_locations, attention_weights):
    """ms_deform_attn.py:159-212.  value [B,S,M,D], spatial_shapes [L,2] (h,w) int64,
    sampling_locations [B,Nq,M,L,P,2] with last dim (x,y) in [0,1], attention_weights [B,Nq,M,L,P]
    -> [B,Nq,M*D]."""
    B, S, M, D = value.shape
    Nq, L, P = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    sizes = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    grids = 2 * sampling_locations - 1  # grid_sample's [-1,1] convention (:168)
    taps, start = [], 0
    for lvl, (h, w) in enumerate(sizes):
        # [B, h*w, M, D] -> [B*M, D, h, w]   (:176-181)
        fmap = value[:, start:start + h * w].flatten(2).transpose(1, 2).reshape(B * M, D, h, w)
        # [B, Nq, M, P, 2] -> [B*M, Nq, P, 2]   (:182-185)
        grid = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        taps.append(F.grid_sample(fmap, grid, mode="bilinear", padding_mode="zeros", align_corners=False))
        start += h * w
    weights = attention_weights.transpose(1, 2).reshape(B * M, 1, Nq, L * P)
    sampled = torch.stack(taps, dim=-2).flatten(-2)  # [B*M, D, Nq, L*P]
    out = (sampled * weights).sum(-1).view(B, M * D, Nq)
    return out.transpose(1, 2).contiguous()


def relation_dim_t(num_pos_feats=16, temperature=10000.0, device=None):
    """position_encoding.py:101-105: temperature ** (2k / num_pos_feats), k < num_pos_feats // 2, fp32."""
    k = torch.arange(num_pos_feats // 2, dtype=torch.float32, device=device)
    return temperature ** (k * 2 / num_pos_feats)


def rel_eager(src_boxes, tgt_boxes, weight, bias, dim_t=None, scale=100.0, eps=1e-5):
    """relation_transformer.py:481-490 + :520-532.  weight [H,64,1,1] (or [H,64]), bias [H]
    -> relu(conv1x1(sincos(box_rel_encoding))) as [B,H,N1,N2].  Geometry carries no gradient."""
    if tgt_boxes is None:
        tgt_boxes = src_boxes
    with torch.no_grad():
        xy1, wh1 = src_boxes.split([2, 2], -1)
        xy2, wh2 = tgt_boxes.split([2, 2], -1)
        dxy = torch.log(torch.abs(xy1.unsqueeze(-2) - xy2.unsqueeze(-3)) / (wh1.unsqueeze(-2) + eps) + 1.0)
        dwh = torch.log((wh1.unsqueeze(-2) + eps) / (wh2.unsqueeze(-3) + eps))
        e = torch.cat([dxy, dwh], -1)  # [B,N1,N2,4]
        if dim_t is None:
            dim_t = relation_dim_t(16, 10000.0, e.device)
        ang = e.unsqueeze(-1) * scale / dim_t.to(e.dtype)  # (x*scale)/dim_t, position_encoding.py:133
        feat = torch.stack((ang.sin(), ang.cos()), dim=-1).flatten(-3)  # [B,N1,N2,64]
        feat = feat.permute(0, 3, 1, 2)
    w4 = weight.reshape(weight.shape[0], -1, 1, 1)
    return F.relu(F.conv2d(feat, w4, bias))
