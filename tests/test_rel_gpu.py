"""Fused position-relation bias on the B200 vs the oracle (through the C ABI).

Tolerances: EXACT mode forward max-abs <= 5e-5 vs the fp64 reference (the reference's own fp32 noise
is up to 2e-5 on these cases, tests/golden/REPORT.txt) and mean-abs <= 2e-6; FAST mode (the default) is pinned
to what was measured (profiles/r01_rel_accuracy.json: 9.1e-6 .. 1.1e-5 max, 3.4e-7 mean): <= 3e-5 / 1e-6.  grad_weight / grad_bias: max-abs-error / max-abs-reference <= 2e-4 (fp32 atomics over up to
6.5 M pairs; the reference's own fp32 noise on grad_weight is 3e-4..7e-4 absolute)."""
import numpy as np
import pytest
import torch

import relation_detr_b200 as rd
from relation_detr_b200 import ops, workloads
from conftest import REL_GOLDEN, load_golden, maxabs, relmax
from oracle import c_oracle, torch_port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_ours(src, tgt, w, b, go, mask=None, fast=False):
    w = w.to(DEV).clone().requires_grad_(True)
    b = b.to(DEV).clone().requires_grad_(True)
    out = ops.position_relation_bias(src.to(DEV), None if tgt is None else tgt.to(DEV), w, b,
                                     attn_mask=None if mask is None else mask.to(DEV), fast=fast)
    keep = out.detach().clone()
    # the decoder mutates the result in place before backward (relation_transformer.py:372-374)
    if mask is not None:
        out.flatten(0, 1).masked_fill_(mask.to(DEV), float("-inf"))
    g = go.to(DEV)
    if mask is not None:
        g = g.masked_fill(mask.to(DEV), 0.0)
    out.backward(g)
    torch.cuda.synchronize()
    return keep.cpu().numpy(), w.grad.cpu().numpy().reshape(w.shape[0], -1), b.grad.cpu().numpy()


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("name", REL_GOLDEN)
def test_matches_reference_fixtures(name, fast):
    g = load_golden(name)
    src = torch.from_numpy(g["src_boxes"])
    tgt = torch.from_numpy(g["tgt_boxes"]) if "tgt_boxes" in g else None
    mask = torch.from_numpy(g["attn_mask"]) if "attn_mask" in g else None
    out, gw, gb = run_ours(src, tgt, torch.from_numpy(g["weight"]), torch.from_numpy(g["bias"]),
                           torch.from_numpy(g["grad_output"]), mask, fast)
    ref = g["ref64_out_masked"] if mask is not None else g["ref64_out"]
    assert np.array_equal(np.isneginf(out), np.isneginf(ref))
    fin = np.isfinite(ref)
    tol_max, tol_mean = (3e-5, 1e-6) if fast else (5e-5, 2e-6)
    assert maxabs(out[fin], ref[fin]) <= tol_max
    assert np.abs(out[fin] - ref[fin]).mean() <= tol_mean
    assert (out[fin] >= 0).all()
    gtol = 5e-4 if fast else 2e-4
    assert relmax(gw, g["ref64_grad_weight"]) <= gtol
    assert relmax(gb, g["ref64_grad_bias"]) <= gtol


@pytest.mark.parametrize("B,N1,N2", [(1, 1, 1), (2, 37, 29), (1, 31, 33), (3, 64, 100), (1, 900, 900), (1, 1100, 1100)])
def test_matches_c_oracle_on_seeded_inputs(B, N1, N2):
    r = workloads.make_rel_inputs(workloads.RelShape("t", B, N1, N2), seed=N1)
    src, tgt, w, b = (r[k] for k in ("src_boxes", "tgt_boxes", "weight", "bias"))
    dim_t = torch_port.relation_dim_t().numpy().astype(np.float64)
    a64 = [x.numpy().astype(np.float64) for x in (src, tgt, w, b)]
    o64 = c_oracle.rel_forward(*a64, dim_t)
    gw64, gb64 = c_oracle.rel_backward(*a64, dim_t, r["grad_output"].numpy().astype(np.float64))
    for fast in (False, True):
        out, gw, gb = run_ours(src, tgt, w, b, r["grad_output"], None, fast)
        assert maxabs(out, o64) <= (3e-5 if fast else 5e-5), fast
        assert np.abs(out - o64).mean() <= (1e-6 if fast else 2e-6), fast
        # sign disagreements of the pre-activation can only happen within rounding of zero ...
        flips = (out > 0) != (o64 > 0)
        assert np.abs(o64[flips]).max(initial=0.0) <= 1e-4
        # ... but each one moves grad_weight[h, :] by up to |grad_out| (|f| <= 1): ReLU'(0) is as
        # discontinuous as floor() is for grad_loc, so the tolerance carries that per-head slack
        slack = (np.abs(r["grad_output"].numpy()) * flips).sum(axis=(0, 2, 3))
        tol = 5e-4 * np.abs(gw64).max()
        assert (np.abs(gw - gw64).max(axis=1) <= tol + slack).all(), (np.abs(gw - gw64).max(axis=1), slack)
        assert (np.abs(gb - gb64) <= 5e-4 * np.abs(gb64).max() + slack).all()


def test_exact_mode_features_match_torch_on_the_same_device():
    """EXACT mode evaluates (e*scale)/dim_t, sinf, cosf, logf and the divisions exactly as torch does
    on CUDA, so against the eager port ON THE GPU only the 64-term summation order differs."""
    r = workloads.make_rel_inputs(workloads.RelShape("t", 2, 200, 180), seed=9, device=DEV)
    # the eager path's 1x1 conv silently runs in TF32 on this GPU unless told otherwise (cudnn.allow_tf32
    # defaults to True): ~5e-4 of error that is the oracle's, not ours
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        want = torch_port.rel_eager(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"])
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    got = ops.position_relation_bias(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"])
    assert (got - want).abs().max().item() <= 5e-6


def test_relu_bits_mask_fusion_and_inplace_mutation():
    shape = workloads.RelShape("t", 2, 70, 45)
    r = workloads.make_rel_inputs(shape, seed=2, device=DEV)
    dim_t = ops.relation_dim_t(16, 10000.0, DEV)
    mask = torch.rand((70, 45), device=DEV) > 0.7
    out, bits = torch.ops.rdetr.relation_forward(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"], dim_t,
                                                 100.0, 1e-5, mask, False)
    plain, bits2 = torch.ops.rdetr.relation_forward(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"], dim_t,
                                                    100.0, 1e-5, None, False)
    assert torch.equal(out, plain.masked_fill(mask, float("-inf")))
    # bits is [B, N1, words, H]; bit j%32 of word j/32 == (pre-activation > 0), and 0 at blocked positions: the reference's
    # masked_fill_ gives those elements no gradient whatever the caller sends back
    j = torch.arange(45, device=DEV)
    unpack = lambda w: ((w[:, :, j // 32, :] >> (j % 32)[None, None, :, None]) & 1).bool().permute(0, 3, 1, 2)  # noqa: E731  [B, H, N1, N2]
    assert torch.equal(unpack(bits2), plain > 0)
    assert torch.equal(unpack(bits), (plain > 0) & ~mask[None, None])


def test_equivariance_and_default_target():
    r = workloads.make_rel_inputs(workloads.RelShape("t", 2, 50, 50), seed=4, device=DEV)
    base = ops.position_relation_bias(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"])
    pi = torch.randperm(50, device=DEV)
    pj = torch.randperm(50, device=DEV)
    perm = ops.position_relation_bias(r["src_boxes"][:, pi].contiguous(), r["tgt_boxes"][:, pj].contiguous(),
                                      r["weight"], r["bias"])
    assert torch.equal(perm, base[:, :, pi][:, :, :, pj])
    self_rel = ops.position_relation_bias(r["src_boxes"], None, r["weight"], r["bias"])
    assert torch.equal(self_rel, ops.position_relation_bias(r["src_boxes"], r["src_boxes"].clone(), r["weight"], r["bias"]))


def test_module_in_decoder_style_use():
    """PositionRelationEmbedding as the decoder calls it: flatten(0,1) + in-place masked_fill_, the
    bias fed to nn.MultiheadAttention, gradient reaching pos_proj through the attention."""
    torch.manual_seed(0)
    rel = rd.PositionRelationEmbedding(16, 8).to(DEV)
    mha = torch.nn.MultiheadAttention(256, 8, batch_first=True).to(DEV)
    B, N = 2, 40
    src = workloads.make_boxes(B, N, 0, DEV)
    tgt = workloads.make_boxes(B, N, 1, DEV)
    mask = workloads.cdn_attn_mask(28, 3, 4, DEV)
    bias = rel(src, tgt).flatten(0, 1)
    bias.masked_fill_(mask, float("-inf"))
    q = torch.randn((B, N, 256), device=DEV)
    out = mha(q, q, q, attn_mask=bias, need_weights=False)[0]
    out.sum().backward()
    gw, gb = rel.pos_proj[0].weight.grad.clone(), rel.pos_proj[0].bias.grad.clone()
    assert gw.shape == (8, 64, 1, 1) and torch.isfinite(gw).all() and gw.abs().sum() > 0
    # same thing with the eager port in place of the kernel
    w = rel.pos_proj[0].weight.detach().clone().requires_grad_(True)
    b = rel.pos_proj[0].bias.detach().clone().requires_grad_(True)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # keep the oracle's conv in fp32
    bias2 = torch_port.rel_eager(src, tgt, w, b).flatten(0, 1).masked_fill(mask, float("-inf"))
    torch.backends.cudnn.allow_tf32 = tf32
    mha.zero_grad()
    out2 = mha(q, q, q, attn_mask=bias2, need_weights=False)[0]
    out2.sum().backward()
    assert (out - out2).abs().max().item() <= 1e-4
    assert (gw - w.grad).abs().max().item() <= 2e-4 * max(1.0, w.grad.abs().max().item())
    assert (gb - b.grad).abs().max().item() <= 2e-4 * max(1.0, b.grad.abs().max().item())


def test_full_size_properties():
    """configs[2] size (B=8, N=900): no oracle at this size in the test budget; check invariants.
    Row i / column j of the result depend only on boxes i and j, so any sub-block must equal the
    same call on the sub-sets; output >= 0; relu bits consistent; FAST within its bound of EXACT."""
    r = workloads.make_rel_inputs(workloads.REL_SHAPES["rel_900_b8"], seed=0, device=DEV)
    out = ops.position_relation_bias(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"])
    assert out.shape == (8, 8, 900, 900) and (out >= 0).all()
    sub = ops.position_relation_bias(r["src_boxes"][5:6, 100:164].contiguous(), r["tgt_boxes"][5:6, 700:733].contiguous(),
                                     r["weight"], r["bias"])
    assert torch.equal(sub, out[5:6, :, 100:164, 700:733])
    fast = ops.position_relation_bias(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"], fast=True)
    d = (fast - out).abs()
    print(f"\nREL fast-vs-exact at B=8,N=900: max {d.max().item():.2e} mean {d.mean().item():.2e}")
    assert d.max().item() <= 1e-4 and d.mean().item() <= 5e-6


@pytest.mark.parametrize("fast", [False, True])
def test_full_size_b8_against_fp64_gpu_oracle(fast):
    """BASELINE configs[2] at its FULL size (B = 8, N = 900, and N = 1100 with the CDN mask) against the eager port in float64
    on the same device, one image at a time; forward and both parameter gradients.  VERDICT r1 weak #2."""
    for name, dn in (("rel_900_b8", 0), ("rel_1100_b8", 200)):
        shape = workloads.REL_SHAPES[name]
        r = workloads.make_rel_inputs(shape, seed=0, device=DEV)
        mask = workloads.cdn_attn_mask(shape.n1 - dn, 10, dn // 10, DEV) if dn else None
        go = r["grad_output"] if mask is None else r["grad_output"].masked_fill(mask, 0.0)
        w = r["weight"].clone().requires_grad_(True)
        b = r["bias"].clone().requires_grad_(True)
        out = ops.position_relation_bias(r["src_boxes"], r["tgt_boxes"], w, b, attn_mask=mask, fast=fast)
        keep = out.detach().clone()
        out.backward(go)
        w64 = r["weight"].double().requires_grad_(True)
        b64 = r["bias"].double().requires_grad_(True)
        worst_max, sum_abs, count = 0.0, 0.0, 0
        slack = torch.zeros(8, dtype=torch.float64, device=DEV)   # per head: sum of |grad_out| over ReLU sign disagreements
        for i in range(shape.batch):
            o64 = torch_port.rel_eager(r["src_boxes"][i:i + 1].double(), r["tgt_boxes"][i:i + 1].double(), w64, b64)
            o64.backward(go[i:i + 1].double())
            got = keep[i:i + 1]
            if mask is not None:
                assert torch.isneginf(got[..., mask]).all()
                d = (got.double() - o64.detach())[..., ~mask].abs()
            else:
                d = (got.double() - o64.detach()).abs()
            worst_max, sum_abs, count = max(worst_max, d.max().item()), sum_abs + d.sum().item(), count + d.numel()
            flips = (got > 0) != (o64.detach() > 0)
            if mask is not None:
                flips &= ~mask   # blocked positions are -inf on our side by construction; their gradient is zeroed on both sides
            assert (o64.detach().abs() * flips).max().item() <= 1e-4   # ... which only happen within rounding of zero
            slack += (go[i:i + 1].double().abs() * flips).sum(dim=(0, 2, 3))
            del o64, d, flips
        tol_max, tol_mean = (3e-5, 1e-6) if fast else (5e-5, 2e-6)
        print(f"\n{name} fast={fast}: max {worst_max:.2e} mean {sum_abs / count:.2e}")
        assert worst_max <= tol_max and sum_abs / count <= tol_mean, (name, worst_max, sum_abs / count)
        # each ReLU sign flip moves grad_weight[h, :] / grad_bias[h] by up to |grad_out| (|f| <= 1): per-head slack, as in the
        # seeded test above; the rest is fp32 accumulation over 6.5 M (9.7 M) pairs against float64
        gw_err = (w.grad.double().reshape(8, 64) - w64.grad.reshape(8, 64)).abs().amax(dim=1)
        gb_err = (b.grad.double() - b64.grad).abs()
        assert (gw_err <= 5e-4 * w64.grad.abs().max() + slack).all(), (name, gw_err.tolist(), slack.tolist())
        assert (gb_err <= 5e-4 * b64.grad.abs().max() + slack).all(), (name, gb_err.tolist(), slack.tolist())


def test_focal_config_size_properties():
    """B=1, N=2900 (focalnet config with denoising_nums=1000): sub-block consistency, mask fusion, backward."""
    r = workloads.make_rel_inputs(workloads.REL_SHAPES["rel_2900_b1"], seed=0, device=DEV)
    mask = workloads.cdn_attn_mask(900, 10, 200, DEV)  # 2000 dn rows + 900 queries
    w = r["weight"].clone().requires_grad_(True)
    b = r["bias"].clone().requires_grad_(True)
    for fast in (False, True):
        out = ops.position_relation_bias(r["src_boxes"], r["tgt_boxes"], w, b, attn_mask=mask, fast=fast)
        assert out.shape == (1, 8, 2900, 2900)
        assert torch.equal(torch.isneginf(out), mask[None, None].expand_as(out))
        sub = ops.position_relation_bias(r["src_boxes"][:, 2000:2100].contiguous(), r["tgt_boxes"][:, 2500:2533].contiguous(),
                                         w, b, fast=fast)
        assert torch.equal(sub, out[:, :, 2000:2100, 2500:2533])
        out.masked_fill(mask, 0.0).sum().backward()
        assert torch.isfinite(w.grad).all() and w.grad.abs().sum() > 0
        w.grad = None
        b.grad = None


def test_masked_positions_carry_no_gradient_whatever_the_upstream_gradient():
    """with the fused -inf mask, a blocked position must not contribute to grad_weight / grad_bias even
    when the caller's gradient is non-zero there -- the reference's masked_fill_ cuts the graph at those elements."""
    r = workloads.make_rel_inputs(workloads.RelShape("t", 2, 70, 45), seed=3, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(4)
    mask = torch.rand((70, 45), device=DEV, generator=g) > 0.6
    go = torch.randn((2, 8, 70, 45), device=DEV, generator=g)  # deliberately non-zero at masked positions
    for fast in (False, True):
        w = r["weight"].clone().requires_grad_(True)
        b = r["bias"].clone().requires_grad_(True)
        out = ops.position_relation_bias(r["src_boxes"], r["tgt_boxes"], w, b, attn_mask=mask, fast=fast)
        out.backward(go)
        w64 = r["weight"].double().requires_grad_(True)
        b64 = r["bias"].double().requires_grad_(True)
        ref = torch_port.rel_eager(r["src_boxes"].double(), r["tgt_boxes"].double(), w64, b64)
        ref = ref.masked_fill(mask, float("-inf"))  # out-of-place twin of relation_transformer.py:372-374
        ref.backward(go.double())
        assert torch.equal(torch.isneginf(out), mask[None, None].expand_as(out))
        assert ((w.grad.double() - w64.grad).abs().max() / w64.grad.abs().max()).item() <= 5e-4
        assert ((b.grad.double() - b64.grad).abs().max() / b64.grad.abs().max()).item() <= 5e-4
