"""Oracle for SURVEY.md section 8 row N1 -- self-attention with the position-relation bias generated on the fly.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  No product kernel exists for this row yet; the oracle
and its fixtures are laid down first so that the kernel has something pinned to be checked against.

What is restated: the decoder layer's self-attention (models/bricks/relation_transformer.py:440-459: the
``nn.MultiheadAttention`` call with ``attn_mask = pos_relation``) together with the producer of that mask
(:369-374: ``PositionRelationEmbedding(src, tgt).flatten(0, 1)`` then ``masked_fill_(attn_mask, -inf)``).
The fused operator's boundary is the attention core between MultiheadAttention's input and output projections:

    out[b,h,i,:] = sum_j softmax_j( q[b,h,i,:].k[b,h,j,:] / sqrt(D) + bias[b,h,i,j] ) v[b,h,j,:]
    bias[b,h,i,j] = relu( W[h,:] . f(src_box[b,i], tgt_box[b,j]) + c[h] ),  -inf where attn_mask[i,j]

with gradients for q, k, v, W, c (none for the boxes, relation_transformer.py:527).  The relation term comes from
the fp64 C oracle (``c_oracle.rel_forward`` / ``rel_backward``, already pinned); the attention arithmetic is plain
numpy in float64.  Pinned against the reference by tests/test_rel_attention_oracle.py on fixtures produced by
oracle/make_golden_rel_attention.py (the reference's own MultiheadAttention + PositionRelationEmbedding, autograd).
"""
from __future__ import annotations

import numpy as np

from . import c_oracle


def forward(q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, attn_mask=None, scale=100.0, eps=1e-5):
    """q, k, v [B,H,N,D]; boxes [B,N,4] cxcywh; weight [H,64]; bias [H]; attn_mask [N,N] bool (True = blocked).
    -> (out [B,H,N,D], probabilities [B,H,N,N], relation bias before masking [B,H,N,N]), all float64."""
    q, k, v = (np.asarray(t, dtype=np.float64) for t in (q, k, v))
    rel = c_oracle.rel_forward(np.asarray(src_boxes, np.float64), np.asarray(tgt_boxes, np.float64),
                               np.asarray(weight, np.float64), np.asarray(bias, np.float64),
                               np.asarray(dim_t, np.float64), scale, eps)
    scores = np.einsum("bhid,bhjd->bhij", q, k) / np.sqrt(q.shape[-1]) + rel
    if attn_mask is not None:
        scores = np.where(np.asarray(attn_mask, bool)[None, None], -np.inf, scores)
    scores = scores - scores.max(-1, keepdims=True)
    p = np.exp(scores)
    p /= p.sum(-1, keepdims=True)
    return np.einsum("bhij,bhjd->bhid", p, v), p, rel


def backward(q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, grad_out, attn_mask=None, scale=100.0, eps=1e-5):
    """-> (grad_q, grad_k, grad_v, grad_weight [H,64], grad_bias [H]).  The gradient of the relation bias is the
    gradient of the pre-softmax scores (masked entries carry probability 0, hence gradient 0); it is pushed
    through the ReLU and the 64 -> H projection by the C oracle's relation backward."""
    q, k, v, go = (np.asarray(t, dtype=np.float64) for t in (q, k, v, grad_out))
    _, p, _ = forward(q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, attn_mask, scale, eps)
    grad_v = np.einsum("bhij,bhid->bhjd", p, go)
    dp = np.einsum("bhid,bhjd->bhij", go, v)
    ds = p * (dp - (dp * p).sum(-1, keepdims=True))
    inv = 1.0 / np.sqrt(q.shape[-1])
    grad_q = np.einsum("bhij,bhjd->bhid", ds, k) * inv
    grad_k = np.einsum("bhij,bhid->bhjd", ds, q) * inv
    grad_w, grad_b = c_oracle.rel_backward(np.asarray(src_boxes, np.float64), np.asarray(tgt_boxes, np.float64),
                                           np.asarray(weight, np.float64), np.asarray(bias, np.float64),
                                           np.asarray(dim_t, np.float64), np.ascontiguousarray(ds), scale, eps)
    return grad_q, grad_k, grad_v, grad_w, grad_b


def split_heads(x, w, b, heads):
    """One third of MultiheadAttention's input projection: x [B,N,E] -> [B,H,N,E/H] (torch/nn/functional.py,
    multi_head_attention_forward: linear, then view(N, B*H, D) per head)."""
    y = np.asarray(x, np.float64) @ np.asarray(w, np.float64).T + np.asarray(b, np.float64)
    B, N, E = y.shape
    return y.reshape(B, N, heads, E // heads).transpose(0, 2, 1, 3)


def merge_heads(x):
    B, H, N, D = x.shape
    return x.transpose(0, 2, 1, 3).reshape(B, N, H * D)
