"""Randomised shapes against the fp64 oracle on the GPU: every supported (L, P, M) corner, ragged query
counts, one-pixel levels, locations far outside the map, with and without the fused prologue."""
import numpy as np
import pytest
import torch

from relation_detr_b200 import ops, workloads
from oracle import torch_port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def test_msda_random_configurations():
    rng = np.random.default_rng(123)
    for case in range(40):
        L, P, M = int(rng.integers(1, 9)), int(rng.integers(1, 9)), int(rng.integers(1, 10))
        levels = tuple((int(rng.integers(1, 20)), int(rng.integers(1, 24))) for _ in range(L))
        B, Nq = int(rng.integers(1, 4)), int(rng.integers(1, 120))
        ss, lsi = workloads.shape_tensors(levels, DEV)
        S = int(ss.prod(1).sum())
        g = torch.Generator(device=DEV).manual_seed(case)
        value = torch.randn((B, S, M, 32), device=DEV, generator=g)
        spread = float(rng.choice([0.2, 1.0, 3.0]))
        loc = (torch.rand((B, Nq, M, L, P, 2), device=DEV, generator=g) - 0.5) * (1 + spread) + 0.5
        attn = torch.rand((B, Nq, M, L, P), device=DEV, generator=g)
        go = torch.randn((B, Nq, M * 32), device=DEV, generator=g)
        v64 = value.double().requires_grad_(True)
        l64 = loc.double().requires_grad_(True)
        a64 = attn.double().requires_grad_(True)
        want = torch_port.msda_grid_sample(v64, ss, l64, a64)
        want.backward(go.double())
        out = ops.msda_forward(value, ss, lsi, loc, attn)
        gv, gl, ga = ops.msda_backward(value, ss, lsi, loc, attn, go)
        tag = f"case {case}: L={L} P={P} M={M} B={B} Nq={Nq} levels={levels}"
        assert (out.double() - want).abs().max().item() <= 2e-5, tag
        assert _rel(gv, v64.grad) <= 1e-4 and _rel(ga, a64.grad) <= 1e-4, tag
        bad = ((gl.double() - l64.grad).abs() > 1e-4 * l64.grad.abs().max().clamp(min=1e-30)).double().mean().item()
        assert bad <= 5e-3, tag


def test_fused_prologue_random_configurations():
    rng = np.random.default_rng(321)
    for case in range(24):
        L, P, M = int(rng.integers(1, 7)), int(rng.integers(1, 7)), int(rng.integers(1, 9))
        levels = tuple((int(rng.integers(1, 16)), int(rng.integers(1, 20))) for _ in range(L))
        B, Nq, R = int(rng.integers(1, 3)), int(rng.integers(1, 90)), int(rng.choice([2, 4]))
        ss, lsi = workloads.shape_tensors(levels, DEV)
        S = int(ss.prod(1).sum())
        g = torch.Generator(device=DEV).manual_seed(1000 + case)
        value = torch.randn((B, S, M, 32), device=DEV, generator=g)
        off = torch.randn((B, Nq, M, L, P, 2), device=DEV, generator=g) * 2.5
        z = torch.randn((B, Nq, M, L * P), device=DEV, generator=g) * 3
        ref = torch.rand((B, Nq, L, R), device=DEV, generator=g)
        if R == 4:
            ref[..., 2:] = ref[..., 2:] * 0.5 + 0.01
        mask = (torch.rand((B, S), device=DEV, generator=g) > 0.7) if case % 2 else None
        go = torch.randn((B, Nq, M * 32), device=DEV, generator=g)
        v = value.clone().requires_grad_(True)
        o = off.clone().requires_grad_(True)
        zz = z.clone().requires_grad_(True)
        out = ops.ms_deform_attn_fused(v, ss, lsi, ref, o, zz, mask)
        out.backward(go)
        v64 = value.double().requires_grad_(True)
        o64 = off.double().requires_grad_(True)
        z64 = z.double().requires_grad_(True)
        vm = v64 if mask is None else v64.masked_fill(mask[..., None, None], 0.0)
        a = z64.softmax(-1).view(B, Nq, M, L, P)
        r64 = ref.double()
        if R == 2:
            wh = torch.stack([ss[..., 1], ss[..., 0]], -1).double()
            loc = r64[:, :, None, :, None, :] + o64 / wh[None, None, None, :, None, :]
        else:
            loc = r64[:, :, None, :, None, :2] + o64 / P * r64[:, :, None, :, None, 2:] * 0.5
        want = torch_port.msda_grid_sample(vm, ss, loc, a)
        want.backward(go.double())
        tag = f"case {case}: L={L} P={P} M={M} B={B} Nq={Nq} R={R} mask={mask is not None}"
        assert (out.double() - want).abs().max().item() <= 2e-5, tag
        assert _rel(v.grad, v64.grad) <= 1e-4 and _rel(zz.grad, z64.grad) <= 2e-4, tag
        bad = ((o.grad.double() - o64.grad).abs() > 1e-4 * o64.grad.abs().max().clamp(min=1e-30)).double().mean().item()
        assert bad <= 5e-3, tag


def test_relation_random_configurations():
    rng = np.random.default_rng(7)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for case in range(16):
            B, N1, N2 = int(rng.integers(1, 4)), int(rng.integers(1, 200)), int(rng.integers(1, 200))
            r = workloads.make_rel_inputs(workloads.RelShape("t", B, N1, N2), seed=case, device=DEV)
            mask = (torch.rand((N1, N2), device=DEV) > 0.6) if case % 2 else None
            w64 = r["weight"].double().requires_grad_(True)
            b64 = r["bias"].double().requires_grad_(True)
            want = torch_port.rel_eager(r["src_boxes"].double(), r["tgt_boxes"].double(), w64, b64,
                                        torch_port.relation_dim_t(device=DEV).double())
            gomask = r["grad_output"] if mask is None else r["grad_output"].masked_fill(mask, 0.0)
            want.backward(gomask.double())
            for fast in (False, True):
                w = r["weight"].clone().requires_grad_(True)
                b = r["bias"].clone().requires_grad_(True)
                out = ops.position_relation_bias(r["src_boxes"], r["tgt_boxes"], w, b, attn_mask=mask, fast=fast)
                fin = torch.isfinite(out)
                if mask is not None:
                    assert torch.equal(~fin, mask[None, None].expand_as(out))
                assert (out[fin].double() - want[fin]).abs().max().item() <= 1e-4, (case, fast)
                out.backward(gomask)
                flips = ((out > 0) != (want > 0)) & fin
                slack = (gomask.abs() * flips).sum(dim=(0, 2, 3)).double()
                tol = 5e-4 * w64.grad.abs().max()
                assert ((w.grad.double() - w64.grad).abs().amax(dim=1) <= tol + slack).all(), (case, fast)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
