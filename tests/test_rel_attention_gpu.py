"""Fused relation attention (csrc/rel_attn.cu) on the B200.

Checked against (a) the fixtures the reference's own MultiheadAttention + PositionRelationEmbedding produced in float64
(oracle/make_golden_rel_attention.py -> tests/golden/relattn_*.npz) through the pinned numpy oracle, and (b) the unfused
pipeline on the same device in float64 (eager relation embedding -> masked_fill -> scaled_dot_product_attention) at the
decoder's real sizes, with and without the denoising mask.  Tolerances: outputs <= 2e-5 max-abs (|out| ~ 1), q/k/v
gradients <= 1e-4 relative to max-abs (north star); grad_weight / grad_bias <= 3e-3: the ReLU of the relation bias is recomputed
in fp32 FAST arithmetic, and a pre-activation within ~1e-5 of zero flips its gate against the float64 reference (the same per-head
slack tests/test_rel_gpu.py carries), each flip moving a weight gradient by one |dS| out of millions of terms.
"""
import os

import numpy as np
import pytest
import torch

from relation_detr_b200 import ops, workloads
from oracle import rel_attention as ra
from oracle import torch_port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    b = torch.as_tensor(b)
    return ((torch.as_tensor(a).double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp(min=1e-30)).item()


@pytest.mark.parametrize("name", ["relattn_plain", "relattn_cdn"])
def test_matches_the_reference_fixtures_through_the_oracle(name):
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
    core, _, _ = ra.forward(q, k, v, z["src_boxes"], z["tgt_boxes"], z["rel_weight"], z["rel_bias"], dim_t, mask)
    g_core = np.random.default_rng(0).standard_normal(core.shape)
    gq, gk, gv, gw, gb = ra.backward(q, k, v, z["src_boxes"], z["tgt_boxes"], z["rel_weight"], z["rel_bias"], dim_t, g_core, mask)

    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=DEV)  # noqa: E731
    tq, tk, tv = (t(a).requires_grad_(True) for a in (q, k, v))
    tw, tb = t(z["rel_weight"]).requires_grad_(True), t(z["rel_bias"]).requires_grad_(True)
    tm = torch.tensor(mask, dtype=torch.bool, device=DEV) if mask is not None else None
    out = ops.relation_attention(tq, tk, tv, t(z["src_boxes"]), t(z["tgt_boxes"]), tw, tb, attn_mask=tm)
    out.backward(t(g_core))
    assert np.abs(out.detach().cpu().numpy() - core).max() <= 2e-5
    assert _rel(tq.grad, gq) <= 1e-4 and _rel(tk.grad, gk) <= 1e-4 and _rel(tv.grad, gv) <= 1e-4
    assert _rel(tw.grad, gw) <= 3e-3 and _rel(tb.grad, gb) <= 3e-3


def _unfused_fp64(q, k, v, src, tgt, w, b, mask):
    """The reference's chain in float64 on the GPU: eager PositionRelationEmbedding -> masked_fill -> SDPA."""
    q64, k64, v64 = (x.detach().double().requires_grad_(True) for x in (q, k, v))
    w64, b64 = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    bias = torch_port.rel_eager(src.double(), tgt.double(), w64, b64)
    if mask is not None:
        bias = bias.masked_fill(mask, float("-inf"))
    out = torch.nn.functional.scaled_dot_product_attention(q64, k64, v64, attn_mask=bias)
    return out, (q64, k64, v64, w64, b64)


@pytest.mark.parametrize("B,N,masked", [(2, 900, False), (2, 1100, True), (1, 333, True), (3, 70, False)])
def test_matches_the_unfused_pipeline_at_decoder_sizes(B, N, masked):
    g = torch.Generator(device=DEV).manual_seed(N)
    q, k, v = (torch.randn((B, 8, N, 32), device=DEV, generator=g) for _ in range(3))
    src = workloads.make_boxes(B, N, 1, DEV)
    tgt = workloads.make_boxes(B, N, 2, DEV)
    w, b = workloads.make_rel_params(8, 64, 0, DEV)
    mask = None
    if masked:
        dn = 200 if N >= 900 else 40
        mask = workloads.cdn_attn_mask(N - dn, 10, dn // 10, DEV)
    go = torch.randn((B, 8, N, 32), device=DEV, generator=g)
    want, leaves = _unfused_fp64(q, k, v, src, tgt, w, b, mask)
    want.backward(go.double())
    tq, tk, tv, tw, tb = (x.clone().requires_grad_(True) for x in (q, k, v, w, b))
    out = ops.relation_attention(tq, tk, tv, src, tgt, tw, tb, attn_mask=mask)
    out.backward(go)
    assert (out.double() - want).abs().max().item() <= 2e-5
    for got, ref, tol, name in ((tq.grad, leaves[0].grad, 1e-4, "grad_q"), (tk.grad, leaves[1].grad, 1e-4, "grad_k"),
                                (tv.grad, leaves[2].grad, 1e-4, "grad_v"), (tw.grad, leaves[3].grad, 3e-3, "grad_weight"),
                                (tb.grad, leaves[4].grad, 3e-3, "grad_bias")):
        assert _rel(got, ref) <= tol, (name, _rel(got, ref))


@pytest.mark.parametrize("B,N,dn", [(1, 70, 40), (2, 333, 40), (1, 1100, 200)])
def test_key_splits_agree_with_the_unsplit_kernel(B, N, dn, monkeypatch):
    """Small grids split the keys over up to 4 CTAs per row block (forward: partial (o, m, l) merged by relattn_combine_kernel;
    backward: dq by reductions).  Every split count must give the unsplit result up to summation order -- including rows for which
    a whole split is blocked by the denoising mask (N = 70, dn = 40: keys 0..31 are one split, blocked for rows >= 40)."""
    g = torch.Generator(device=DEV).manual_seed(N + dn)
    q, k, v = (torch.randn((B, 8, N, 32), device=DEV, generator=g) for _ in range(3))
    src, tgt = workloads.make_boxes(B, N, 1, DEV), workloads.make_boxes(B, N, 2, DEV)
    w, b = workloads.make_rel_params(8, 64, 0, DEV)
    mask = workloads.cdn_attn_mask(N - dn, 10, dn // 10, DEV)
    go = torch.randn((B, 8, N, 32), device=DEV, generator=g)
    runs = {}
    for splits in (1, 2, 3, 4):
        monkeypatch.setenv("RDETR_RELATTN_SPLITS", str(splits))
        tq, tk, tv, tw, tb = (x.clone().requires_grad_(True) for x in (q, k, v, w, b))
        out = ops.relation_attention(tq, tk, tv, src, tgt, tw, tb, attn_mask=mask)
        out.backward(go)
        runs[splits] = (out.detach(), tq.grad, tk.grad, tv.grad, tw.grad, tb.grad)
    monkeypatch.delenv("RDETR_RELATTN_SPLITS")
    assert torch.isfinite(runs[1][0]).all()
    for splits in (2, 3, 4):
        assert (runs[splits][0] - runs[1][0]).abs().max().item() <= 2e-6, splits
        for got, ref, name in zip(runs[splits][1:], runs[1][1:], ("grad_q", "grad_k", "grad_v", "grad_weight", "grad_bias")):
            assert _rel(got, ref) <= 2e-5, (splits, name, _rel(got, ref))


def test_opcheck_and_errors():
    g = torch.Generator(device=DEV).manual_seed(0)
    q, k, v = (torch.randn((1, 8, 40, 32), device=DEV, generator=g).requires_grad_(True) for _ in range(3))
    src, tgt = workloads.make_boxes(1, 40, 1, DEV), workloads.make_boxes(1, 40, 2, DEV)
    w, b = workloads.make_rel_params(8, 64, 0, DEV)
    dim_t = ops.relation_dim_t(16, 10000.0, DEV)
    args = (q, k, v, src, tgt, w.requires_grad_(True), b.requires_grad_(True), dim_t, 100.0, 1e-5, None)
    torch.library.opcheck(torch.ops.rdetr.relation_attention_forward.default, args,
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    with pytest.raises(RuntimeError, match="H=4"):
        ops.relation_attention(q[:, :4].contiguous(), k[:, :4].contiguous(), v[:, :4].contiguous(), src, tgt, w[:4].contiguous(), b[:4].contiguous())
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.relation_attention(q.cpu(), k.cpu(), v.cpu(), src.cpu(), tgt.cpu(), w.cpu(), b.cpu())
