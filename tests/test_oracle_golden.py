"""The oracle is pinned here: C restatement and torch port vs fixtures generated FROM THE REFERENCE
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import MSDA_GOLDEN, REL_GOLDEN, load_golden, maxabs, relmax
from oracle import c_oracle, ref_import, torch_port


@pytest.mark.parametrize("name", MSDA_GOLDEN)
def test_c_oracle_msda_f64_matches_reference(name):
    g = load_golden(name)
    a = [g["value"].astype(np.float64), g["spatial_shapes"], g["level_start_index"],
         g["sampling_locations"].astype(np.float64), g["attention_weights"].astype(np.float64)]
    out = c_oracle.msda_forward(*a)
    gv, gl, ga = c_oracle.msda_backward(*a, g["grad_output"].astype(np.float64))
    assert maxabs(out, g["ref64_out"]) < 1e-12
    assert maxabs(gv, g["ref64_grad_value"]) < 1e-12
    assert maxabs(gl, g["ref64_grad_loc"]) < 1e-10
    assert maxabs(ga, g["ref64_grad_attn"]) < 1e-11


@pytest.mark.parametrize("name", MSDA_GOLDEN)
def test_c_oracle_msda_f32_within_noise_floor(name):
    g = load_golden(name)
    a = [g["value"], g["spatial_shapes"], g["level_start_index"], g["sampling_locations"], g["attention_weights"]]
    out = c_oracle.msda_forward(*a)
    gv, gl, ga = c_oracle.msda_backward(*a, g["grad_output"])
    assert out.dtype == np.float32
    # fp32 forward within 1e-5 abs of the fp64 reference (north star), grads within 1e-4 rel
    assert maxabs(out, g["ref64_out"]) < 1e-5
    assert relmax(gv, g["ref64_grad_value"]) < 1e-4
    assert relmax(ga, g["ref64_grad_attn"]) < 1e-4
    if "strict" in name:  # grad_loc is only comparable away from pixel boundaries
        assert relmax(gl, g["ref64_grad_loc"]) < 1e-4


@pytest.mark.parametrize("name", MSDA_GOLDEN)
def test_torch_port_msda_is_the_reference_computation(name):
    g = load_golden(name)
    v = torch.from_numpy(g["value"]).requires_grad_(True)
    loc = torch.from_numpy(g["sampling_locations"]).requires_grad_(True)
    attn = torch.from_numpy(g["attention_weights"]).requires_grad_(True)
    out = torch_port.msda_grid_sample(v, torch.from_numpy(g["spatial_shapes"]), loc, attn)
    out.backward(torch.from_numpy(g["grad_output"]))
    # same ATen calls in the same order as the reference: allow only last-bit noise across builds
    assert maxabs(out.detach().numpy(), g["ref32_out"]) < 1e-6
    assert relmax(v.grad.numpy(), g["ref32_grad_value"]) < 1e-6
    assert relmax(attn.grad.numpy(), g["ref32_grad_attn"]) < 1e-6
    assert relmax(loc.grad.numpy(), g["ref32_grad_loc"]) < 1e-5


@pytest.mark.parametrize("name", REL_GOLDEN)
def test_c_oracle_rel_matches_reference(name):
    g = load_golden(name)
    src = g["src_boxes"]
    tgt = g.get("tgt_boxes", src)
    d64 = g["dim_t"].astype(np.float64)
    out = c_oracle.rel_forward(src.astype(np.float64), tgt.astype(np.float64), g["weight"].astype(np.float64),
                               g["bias"].astype(np.float64), d64)
    assert maxabs(out, g["ref64_out"]) < 1e-12
    go = g["grad_output"].astype(np.float64)
    if "attn_mask" in g:
        masked = c_oracle.rel_forward(src.astype(np.float64), tgt.astype(np.float64), g["weight"].astype(np.float64),
                                      g["bias"].astype(np.float64), d64, attn_mask=g["attn_mask"])
        assert np.array_equal(np.isneginf(masked), np.isneginf(g["ref64_out_masked"]))
        fin = np.isfinite(masked)
        assert maxabs(masked[fin], g["ref64_out_masked"][fin]) < 1e-12
        go = np.where(g["attn_mask"][None, None], 0.0, go)
    gw, gb = c_oracle.rel_backward(src.astype(np.float64), tgt.astype(np.float64), g["weight"].astype(np.float64),
                                   g["bias"].astype(np.float64), d64, go)
    assert maxabs(gw, g["ref64_grad_weight"]) < 1e-10
    assert maxabs(gb, g["ref64_grad_bias"]) < 1e-10
    # fp32 flavour: within the reference's own fp32 noise (REPORT.txt: <= 2e-5 on these cases)
    out32 = c_oracle.rel_forward(src, tgt, g["weight"], g["bias"], g["dim_t"])
    assert maxabs(out32, g["ref64_out"]) < 5e-5


@pytest.mark.parametrize("name", REL_GOLDEN)
def test_torch_port_rel_is_the_reference_computation(name):
    g = load_golden(name)
    src = torch.from_numpy(g["src_boxes"])
    tgt = torch.from_numpy(g["tgt_boxes"]) if "tgt_boxes" in g else None
    w = torch.from_numpy(g["weight"]).requires_grad_(True)
    b = torch.from_numpy(g["bias"]).requires_grad_(True)
    out = torch_port.rel_eager(src, tgt, w, b)
    go = torch.from_numpy(g["grad_output"])
    if "attn_mask" in g:
        go = go.masked_fill(torch.from_numpy(g["attn_mask"]), 0.0)
    out.backward(go)
    assert maxabs(out.detach().numpy(), g["ref32_out"]) < 1e-6
    assert relmax(w.grad.numpy(), g["ref32_grad_weight"]) < 1e-5
    assert relmax(b.grad.numpy(), g["ref32_grad_bias"]) < 1e-5


def test_oracle_edge_cases():
    # empty query set, and a sample exactly on the validity boundary contributes nothing (cuh:274-277)
    shapes = np.array([[2, 3]], dtype=np.int64)
    lsi = np.array([0], dtype=np.int64)
    value = np.ones((1, 6, 1, 32), dtype=np.float64)
    loc = np.zeros((1, 0, 1, 1, 1, 2), dtype=np.float64)
    attn = np.zeros((1, 0, 1, 1, 1), dtype=np.float64)
    assert c_oracle.msda_forward(value, shapes, lsi, loc, attn).shape == (1, 0, 32)
    loc = np.array([-0.5 / 3, 0.5], dtype=np.float64).reshape(1, 1, 1, 1, 1, 2)  # w_im == -1 exactly
    attn = np.ones((1, 1, 1, 1, 1), dtype=np.float64)
    assert np.all(c_oracle.msda_forward(value, shapes, lsi, loc, attn) == 0)
    loc = np.array([0.5, 0.5], dtype=np.float64).reshape(1, 1, 1, 1, 1, 2)  # interior: average of ones
    assert np.allclose(c_oracle.msda_forward(value, shapes, lsi, loc, attn), 1.0)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted (GPU box)")
def test_fixtures_are_reproducible_from_the_reference():
    """Re-run the reference on a stored input and compare with the stored output."""
    msda_ref, PRE, _, _ = ref_import.load()
    g = load_golden("msda_tiny_oob")
    out = msda_ref(torch.from_numpy(g["value"]), torch.from_numpy(g["spatial_shapes"]),
                   torch.from_numpy(g["sampling_locations"]), torch.from_numpy(g["attention_weights"]))
    assert maxabs(out.numpy(), g["ref32_out"]) < 1e-6
    g = load_golden("rel_tiny")
    mod = PRE(16, 8)
    with torch.no_grad():
        mod.pos_proj[0].weight.copy_(torch.from_numpy(g["weight"]).view(8, 64, 1, 1))
        mod.pos_proj[0].bias.copy_(torch.from_numpy(g["bias"]))
        out = mod(torch.from_numpy(g["src_boxes"]), torch.from_numpy(g["tgt_boxes"]))
    assert maxabs(out.numpy(), g["ref32_out"]) < 1e-6


def test_two_restatements_agree_on_random_shapes():
    """C restatement vs torch port (both pinned to the fixtures above) on 12 random configurations in
    fp64, including D != 32, one-pixel levels and far out-of-range samples: <= 1e-11 everywhere."""
    rng = np.random.default_rng(0)
    for case in range(12):
        L = int(rng.integers(1, 6))
        levels = tuple((int(rng.integers(1, 12)), int(rng.integers(1, 14))) for _ in range(L))
        B, M, D = int(rng.integers(1, 3)), int(rng.integers(1, 5)), int(rng.choice([4, 8, 32]))
        P, Nq = int(rng.integers(1, 5)), int(rng.integers(1, 30))
        S = sum(h * w for h, w in levels)
        value = rng.standard_normal((B, S, M, D))
        loc = rng.uniform(-0.6, 1.6, (B, Nq, M, L, P, 2))
        attn = rng.uniform(0, 1, (B, Nq, M, L, P))
        go = rng.standard_normal((B, Nq, M * D))
        shapes = np.array(levels, dtype=np.int64)
        lsi = np.concatenate([[0], np.cumsum(shapes.prod(1))[:-1]]).astype(np.int64)
        out = c_oracle.msda_forward(value, shapes, lsi, loc, attn)
        gv, gl, ga = c_oracle.msda_backward(value, shapes, lsi, loc, attn, go)
        tv = torch.from_numpy(value).requires_grad_(True)
        tl = torch.from_numpy(loc).requires_grad_(True)
        ta = torch.from_numpy(attn).requires_grad_(True)
        tout = torch_port.msda_grid_sample(tv, torch.from_numpy(shapes), tl, ta)
        tout.backward(torch.from_numpy(go))
        assert maxabs(out, tout.detach().numpy()) < 1e-11, case
        assert maxabs(gv, tv.grad.numpy()) < 1e-11 and maxabs(ga, ta.grad.numpy()) < 1e-11, case
        assert maxabs(gl, tl.grad.numpy()) < 1e-9, case
    for case in range(6):
        B, N1, N2 = int(rng.integers(1, 3)), int(rng.integers(1, 40)), int(rng.integers(1, 40))
        src = np.concatenate([rng.uniform(0, 1, (B, N1, 2)), rng.uniform(1e-3, 0.5, (B, N1, 2))], -1)
        tgt = np.concatenate([rng.uniform(0, 1, (B, N2, 2)), rng.uniform(1e-3, 0.5, (B, N2, 2))], -1)
        w, b = rng.uniform(-0.125, 0.125, (8, 64)), rng.uniform(-0.125, 0.125, 8)
        dim_t = torch_port.relation_dim_t().numpy().astype(np.float64)
        out = c_oracle.rel_forward(src, tgt, w, b, dim_t)
        tw = torch.from_numpy(w).requires_grad_(True)
        tb = torch.from_numpy(b).requires_grad_(True)
        tout = torch_port.rel_eager(torch.from_numpy(src), torch.from_numpy(tgt), tw, tb, torch.from_numpy(dim_t))
        go = rng.standard_normal(out.shape)
        tout.backward(torch.from_numpy(go))
        gw, gb = c_oracle.rel_backward(src, tgt, w, b, dim_t, go)
        assert maxabs(out, tout.detach().numpy()) < 1e-11, case
        assert maxabs(gw, tw.grad.numpy()) < 1e-9 and maxabs(gb, tb.grad.numpy()) < 1e-9, case
