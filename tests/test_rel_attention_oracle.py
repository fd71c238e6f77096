"""CPU: the oracle for SURVEY §8 row N1 (self-attention with the relation bias generated on the fly) against
fixtures produced by the reference's own MultiheadAttention + PositionRelationEmbedding in float64
(oracle/make_golden_rel_attention.py).  No product kernel exists for this row yet: this pins the checker first."""
import os

import numpy as np
import pytest

from oracle import rel_attention as ra
from oracle import torch_port

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["relattn_plain", "relattn_cdn"])
def test_oracle_reproduces_the_reference_self_attention(name):
    z = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    z.update({k: v.astype(np.float64) for k, v in np.load(os.path.join(GOLDEN, "relattn_weights.npz")).items()})
    H, E = 8, 256
    w, b = z["in_proj_weight"], z["in_proj_bias"]
    qp = z["query"] + z["query_pos"]
    q = ra.split_heads(qp, w[:E], b[:E], H)
    k = ra.split_heads(qp, w[E:2 * E], b[E:2 * E], H)
    v = ra.split_heads(z["query"], w[2 * E:], b[2 * E:], H)
    mask = z["attn_mask"] if z["attn_mask"].size else None
    dim_t = torch_port.relation_dim_t().double().numpy()
    core, p, rel = ra.forward(q, k, v, z["src_boxes"], z["tgt_boxes"], z["rel_weight"], z["rel_bias"], dim_t, mask)
    out = ra.merge_heads(core) @ z["out_proj_weight"].T + z["out_proj_bias"]
    assert np.abs(out - z["out"]).max() <= 1e-12
    if mask is not None:
        assert (p[:, :, mask] == 0).all()
    assert (rel >= 0).all()

    # gradients: chain the stored upstream gradient through the output projection by hand
    g_core = (z["grad_out"] @ z["out_proj_weight"]).reshape(core.shape[0], core.shape[2], H, E // H).transpose(0, 2, 1, 3)
    gq, gk, gv, gw, gb = ra.backward(q, k, v, z["src_boxes"], z["tgt_boxes"], z["rel_weight"], z["rel_bias"], dim_t, g_core, mask)
    assert np.abs(gw - z["grad_rel_weight"]).max() <= 1e-11 * max(1.0, np.abs(z["grad_rel_weight"]).max())
    assert np.abs(gb - z["grad_rel_bias"]).max() <= 1e-11 * max(1.0, np.abs(z["grad_rel_bias"]).max())
    # d in_proj_bias = [gq ; gk ; gv] summed over batch and positions
    gb_in = np.concatenate([ra.merge_heads(t).sum((0, 1)) for t in (gq, gk, gv)])
    assert np.abs(gb_in - z["grad_in_proj_bias"]).max() <= 1e-10 * max(1.0, np.abs(z["grad_in_proj_bias"]).max())
    g_query = (ra.merge_heads(gq) @ w[:E] + ra.merge_heads(gk) @ w[E:2 * E] + ra.merge_heads(gv) @ w[2 * E:])
    assert np.abs(g_query - z["grad_query"]).max() <= 1e-10 * max(1.0, np.abs(z["grad_query"]).max())
