"""memory_fusion input Linear as a K-split tcgen05 GEMM (csrc/memfuse.cu) on the B200.

Reference = what upstream computes (relation_transformer.py:168-173, 203-204): relu(cat(queries, -1) @ W^T + b), in float64.
The kernel multiplies in TF32 (10-bit mantissa products, fp32 accumulation), so the bound is TF32's: max-abs error <= 2e-3 of the
largest output for K = 1792 unit-variance terms -- the bf16 GEMM autocast runs upstream is ~8x looser.  Gradients are plain
library GEMMs on the unconcatenated tensors and must match autograd through the reference expression.
"""
import pytest
import torch

from relation_detr_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(lead, nsrc, C=256, N=256, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    srcs = [torch.randn((*lead, C), device=DEV, generator=g) for _ in range(nsrc)]
    w = torch.randn((N, nsrc * C), device=DEV, generator=g) / (nsrc * C) ** 0.5
    b = torch.randn((N,), device=DEV, generator=g) * 0.1
    return srcs, w, b


@pytest.mark.parametrize("lead,nsrc", [((2, 1000), 7), ((1, 128), 1), ((3, 77), 7), ((1, 5), 2), ((2, 22323), 7)])
@pytest.mark.parametrize("relu", [True, False])
def test_forward_matches_cat_linear(lead, nsrc, relu):
    srcs, w, b = _case(lead, nsrc, seed=nsrc)
    out = ops.memory_fusion_linear(srcs, w, b, relu)
    want = torch.cat([s.double() for s in srcs], -1) @ w.double().t() + b.double()
    if relu:
        want = want.relu()
    assert out.shape == want.shape and out.dtype == torch.float32
    err = (out.double() - want).abs().max().item()
    assert err <= 2e-3 * want.abs().max().item(), err
    # and it is not merely "close": TF32 keeps 10 mantissa bits, bf16 would sit near 1e-2
    assert (out.double() - want).abs().mean().item() <= 3e-4 * want.abs().max().item()


def test_gradients_match_autograd_through_the_reference_expression():
    srcs, w, b = _case((2, 300), 7, seed=3)
    go = torch.randn((2, 300, 256), device=DEV)
    xs = [s.clone().requires_grad_(True) for s in srcs]
    wt, bt = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = ops.memory_fusion_linear(xs, wt, bt, True)
    out.backward(go)
    xr = [s.double().requires_grad_(True) for s in srcs]
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = (torch.cat(xr, -1) @ wr.t() + br).relu()
    # the ReLU gate of the fp32 / TF32 forward: use the kernel's own output sign so that only the GEMMs are compared
    (ref * 0 + (torch.cat(xr, -1) @ wr.t() + br) * (out.detach() > 0)).backward(go.double())
    rel = lambda a, c: ((a.double() - c).abs().max() / c.abs().max()).item()  # noqa: E731
    for x, r in zip(xs, xr):
        assert rel(x.grad, r.grad) <= 1e-3
    assert rel(wt.grad, wr.grad) <= 1e-3 and rel(bt.grad, br.grad) <= 1e-4


def test_rejects_what_it_is_not_built_for():
    srcs, w, b = _case((1, 64), 2)
    with pytest.raises(RuntimeError, match="N = 256"):
        ops.memory_fusion_linear(srcs, w[:128].contiguous(), b[:128].contiguous())
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.memory_fusion_linear([s.cpu() for s in srcs], w.cpu(), b.cpu())
