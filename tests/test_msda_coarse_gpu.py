"""Coarse-level grad_value accumulation in shared memory (csrc/msda_bwd_coarse.cu) on the B200.

Same bar as tests/test_msda_gpu.py (gradients <= 1e-4 relative to max-abs vs the fp64 oracle).  Every case runs the
backward with the coarse kernel forced on (mode 2) against the oracle AND against the scatter-only path (mode 1) on
the same inputs: pyramids whose levels fall into the big class (281..1056 pixels), the small class (<= 280), or
neither; uniform locations (no locality: the four streams of a group collide often, so the one-at-a-time fallback
runs), out-of-range locations, P in {1, 2, 4, 8}, a point count that does not divide 8 (falls back to the scatter),
odd head counts, decoder-sized Nq, the fused prologue with padding masks, bf16, CUDA-graph capture, and BASELINE
configs[1] at full size.
"""
import numpy as np
import pytest
import torch

from relation_detr_b200 import _lib, ops, workloads
from conftest import maxabs, relmax
from oracle import torch_port
from test_msda_gpu import _fused_case, _oracle_pipeline, run_ours, run_torch_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

PYR_BIG = ((50, 84), (25, 42), (13, 21), (7, 11))       # levels 1..3 taken: 1050 (big), 273 and 77 (small)
PYR_EDGE = ((33, 32), (24, 44), (10, 28), (1, 1))        # 1056 = cap exactly, 1056, 280 = small cap exactly, 1 pixel
PYR_NONE = ((40, 30), (36, 30))                          # nothing qualifies (1200, 1080 pixels)


@pytest.fixture(autouse=True)
def _restore_mode():
    yield
    _lib.lib().rdetr_msda_set_coarse_mode(0)


def _mode(mode):
    _lib.check(_lib.lib().rdetr_msda_set_coarse_mode(mode), "set_coarse_mode")


def _both(inp, dtype=torch.float32):
    _mode(1)
    flat = run_ours(inp, dtype)
    _mode(2)
    return flat, run_ours(inp, dtype)


@pytest.mark.parametrize("levels", [PYR_BIG, PYR_EDGE, PYR_NONE])
@pytest.mark.parametrize("loc_kind", ["S", "U", "oob", "strict"])
@pytest.mark.parametrize("nq", [0, 333])
def test_coarse_fp32_matches_oracle_and_scatter(levels, loc_kind, nq):
    shape = workloads.MsdaShape("t", 2, levels, nq)
    inp = workloads.make_msda_inputs(shape, loc_kind, seed=11)
    ref = run_torch_oracle(inp, torch.float64)
    flat, r = _both(inp)
    assert relmax(r["grad_value"], ref["grad_value"]) <= 1e-4
    assert relmax(r["grad_attn"], ref["grad_attn"]) <= 1e-4
    if loc_kind == "strict":
        assert relmax(r["grad_loc"], ref["grad_loc"]) <= 1e-4
    assert relmax(r["grad_value"], flat["grad_value"]) <= 1e-5
    # grad_loc / grad_attn do not go through the coarse kernel at all
    assert np.array_equal(r["grad_attn"], flat["grad_attn"]) and np.array_equal(r["grad_loc"], flat["grad_loc"])


@pytest.mark.parametrize("B,M,P", [(1, 8, 1), (3, 5, 2), (2, 1, 8), (1, 3, 4), (2, 8, 3)])
def test_coarse_points_heads_and_batches(B, M, P):
    shape = workloads.MsdaShape("t", B, PYR_BIG, 0, heads=M, points=P)
    inp = workloads.make_msda_inputs(shape, "oob", seed=B * 7 + M + P)
    ref = run_torch_oracle(inp, torch.float64)
    flat, r = _both(inp)
    assert relmax(r["grad_value"], ref["grad_value"]) <= 1e-4
    assert relmax(r["grad_value"], flat["grad_value"]) <= 1e-5


def test_coarse_all_samples_on_one_pixel():
    # every query samples the same spot of every level: each group of four streams collides (one-at-a-time path), and
    # one row of the plane receives every update
    shape = workloads.MsdaShape("t", 1, PYR_BIG, 0)
    inp = workloads.make_msda_inputs(shape, "S", seed=3)
    inp["sampling_locations"] = torch.full_like(inp["sampling_locations"], 0.37)
    ref = run_torch_oracle(inp, torch.float64)
    flat, r = _both(inp)
    assert relmax(r["grad_value"], ref["grad_value"]) <= 1e-4
    assert relmax(r["grad_value"], flat["grad_value"]) <= 2e-5


def test_coarse_bf16_matches_oracle():
    shape = workloads.MsdaShape("t", 2, PYR_BIG, 0)
    inp = workloads.make_msda_inputs(shape, "S", seed=6)
    inp["value"] = inp["value"].bfloat16().float()
    inp["grad_output"] = inp["grad_output"].bfloat16().float()
    ref = run_torch_oracle(inp, torch.float64)
    _mode(2)
    r = run_ours(inp, torch.bfloat16)
    assert relmax(r["grad_value"], ref["grad_value"]) <= 2e-2
    assert relmax(r["grad_attn"], ref["grad_attn"]) <= 1e-4


@pytest.mark.parametrize("ref_dim,with_mask", [(2, False), (2, True), (4, True)])
def test_coarse_fused_prologue_matches_oracle_and_scatter(ref_dim, with_mask):
    levels = PYR_BIG
    S = sum(h * w for h, w in levels)
    value, ss, lsi, ref, offsets, logits, mask, go = _fused_case(2, S, levels, 8, 4, ref_dim, 9, with_mask)
    if ref_dim == 2:
        ref = workloads.full_reference_points(levels, DEV)[None, :, None, :].expand(2, S, len(levels), 2).contiguous()
    want_out, want_gv, want_go, want_gz = _oracle_pipeline(value, ss, ref, offsets, logits, mask, go)
    got = {}
    for mode in (1, 2):
        _mode(mode)
        v = value.clone().requires_grad_(True)
        off = offsets.clone().requires_grad_(True)
        z = logits.clone().requires_grad_(True)
        out = ops.ms_deform_attn_fused(v, ss, lsi, ref, off, z, mask)
        out.backward(go)
        got[mode] = (out.detach(), v.grad, off.grad, z.grad)
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()
    assert rel(got[2][1], want_gv) <= 1e-4 and rel(got[2][3], want_gz) <= 1e-4
    if mask is not None:
        assert torch.count_nonzero(got[2][1][mask]) == 0
    assert rel(got[1][1], got[2][1].double()) <= 1e-5
    assert torch.equal(got[1][2], got[2][2]) and torch.equal(got[1][3], got[2][3])


def test_coarse_fused_bf16():
    S = sum(h * w for h, w in PYR_BIG)
    value, ss, lsi, ref, offsets, logits, mask, go = _fused_case(1, S, PYR_BIG, 8, 4, 2, 13, True)
    vb, ob_, zb = value.bfloat16(), offsets.bfloat16(), logits.bfloat16()
    want_out, want_gv, want_go, want_gz = _oracle_pipeline(vb.float(), ss, ref, ob_.float(), zb.float(), mask, go.bfloat16().float())
    _mode(2)
    v = vb.clone().requires_grad_(True)
    off = ob_.clone().requires_grad_(True)
    z = zb.clone().requires_grad_(True)
    out = ops.ms_deform_attn_fused(v, ss, lsi, ref, off, z, mask)
    out.backward(go.bfloat16())
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    assert rel(v.grad, want_gv) <= 2e-2 and rel(z.grad, want_gz) <= 2e-2


def test_coarse_under_cuda_graph_capture():
    shape = workloads.MsdaShape("t", 2, PYR_BIG, 0)
    inp = workloads.make_msda_inputs(shape, "S", seed=21, device=DEV)
    args = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"],
            inp["attention_weights"], inp["grad_output"])
    _mode(2)
    eager = ops.msda_backward(*args)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.msda_backward(*args)  # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        captured = ops.msda_backward(*args)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    for a, b in zip(eager, captured):
        assert ((a - b).abs().max() <= 1e-5 * a.abs().max()).item()


def test_coarse_full_size_b8_against_fp64_gpu_oracle():
    """BASELINE configs[1] at its full size (B = 8, S = Nq = 22 323) with the coarse kernel on, against the fp64 oracle
    on the same device, one image at a time for the oracle."""
    shape = workloads.MSDA_SHAPES["msda_enc_800x1333_b8"]
    inp = workloads.make_msda_inputs(shape, "S", seed=0, device=DEV)
    _mode(2)
    gv, gl, ga = ops.msda_backward(inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"],
                                   inp["attention_weights"], inp["grad_output"])
    _mode(1)
    gv1, _, _ = ops.msda_backward(inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"],
                                  inp["attention_weights"], inp["grad_output"])
    assert ((gv - gv1).abs().max() / gv1.abs().max()).item() <= 1e-5
    worst = 0.0
    for b in range(shape.batch):
        one = {k: (t[b:b + 1] if t.dim() > 2 else t) for k, t in inp.items()}
        v64 = one["value"].double().requires_grad_(True)
        o64 = torch_port.msda_grid_sample(v64, one["spatial_shapes"], one["sampling_locations"].double(),
                                          one["attention_weights"].double())
        o64.backward(one["grad_output"].double())
        worst = max(worst, ((gv[b:b + 1].double() - v64.grad).abs().max() / v64.grad.abs().max()).item())
        del v64, o64
    assert worst <= 1e-4, worst
