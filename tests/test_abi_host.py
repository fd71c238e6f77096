"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the custom ops have shape-only fake kernels, nothing falls back to the CPU, and the
drop-in modules keep the reference's parameter names."""
import ctypes
import os
import re

import pytest
import torch

import relation_detr_b200 as rd
from relation_detr_b200 import _lib, build, dist, modules, ops, workloads
from oracle import ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "rdetr_ops.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rdetr_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built_in_tree():
    path = build.build()  # no-op when up to date; cross-compiles for sm_100a without a GPU
    assert os.path.dirname(path) == os.path.join(ROOT, "relation-detr_b200")
    assert os.path.exists(path)


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), (names, _lib.EXPORTED_SYMBOLS)
    handle = ctypes.CDLL(build.build())
    for n in names:
        assert hasattr(handle, n), n
    assert _lib.lib().rdetr_abi_version() == 1


def test_abi_rejects_bad_arguments_without_a_gpu():
    L = _lib.lib()
    # D = 16 is not supported: error code + message, no crash, no launch
    rc = L.rdetr_msda_forward(1, 1, 1, 1, 1, 1, 1, 10, 8, 16, 1, 1, 4, 0, None)
    assert rc == 2 and b"D=16" in L.rdetr_last_error()
    # maximum sizes: S*M*D must fit 31 bits, at most 8 levels / 8 points, dtype 0 or 1
    assert L.rdetr_msda_forward(1, 1, 1, 1, 1, 1, 1, 1 << 23, 8, 32, 4, 1, 4, 0, None) == 2 and b"31 bits" in L.rdetr_last_error()
    assert L.rdetr_msda_forward(1, 1, 1, 1, 1, 1, 1, 10, 8, 32, 9, 1, 4, 0, None) == 2 and b"L=9" in L.rdetr_last_error()
    assert L.rdetr_msda_forward(1, 1, 1, 1, 1, 1, 1, 10, 8, 32, 4, 1, 4, 7, None) == 2 and b"value_dtype 7" in L.rdetr_last_error()
    assert L.rdetr_msda_fused_forward(1, 1, 1, 1, 1, 1, None, 1, 1, 10, 8, 32, 4, 1, 4, 3, 0, None) == 1 and b"2 or 4" in L.rdetr_last_error()
    # empty batches / query sets are a no-op, not an error
    assert L.rdetr_msda_forward(None, None, None, None, None, None, 0, 10, 8, 32, 4, 5, 4, 0, None) == 0
    assert L.rdetr_msda_forward(None, None, None, None, None, None, 2, 10, 8, 32, 4, 0, 4, 0, None) == 0
    assert L.rdetr_relation_forward(None, None, None, None, None, 100.0, 1e-5, None, None, None, 0, 4, 4, 8, 0, None, 0, None) == 0
    rc = L.rdetr_msda_forward(None, None, None, None, None, None, 1, 10, 8, 32, 1, 1, 4, 0, None)
    assert rc == 1 and b"null" in L.rdetr_last_error()
    rc = L.rdetr_relation_forward(None, None, None, None, None, 100.0, 1e-5, None, None, None, 1, 4, 4, 6, 0, None, 0, None)
    assert rc == 2 and b"H=6" in L.rdetr_last_error()
    assert L.rdetr_msda_backward_workspace_bytes(2, 10, 8, 32, 1, 1, 4, 1) == 2 * 10 * 8 * 32 * 4
    assert L.rdetr_msda_backward_workspace_bytes(2, 10, 8, 32, 1, 1, 4, 0) == 0
    assert L.rdetr_relation_workspace_bytes(2, 10, 6, 1) == 2 * 16 * 36 * 4 and L.rdetr_relation_workspace_bytes(2, 10, 6, 0) == 0
    # assignment solver / cost kernel: extents are validated before anything touches a device
    import ctypes
    one = lambda v: (ctypes.c_int64 * 1)(v)  # noqa: E731
    ptr = (ctypes.c_void_p * 1)(16)
    assert L.rdetr_lsap_workspace_bytes(one(900), one(40), 1) == (900 * 40 * 4 + 255) // 256 * 256
    assert L.rdetr_lsap_workspace_bytes(one(40), one(900), 1) == 0
    assert L.rdetr_lsap_solve(ptr, one(3), one(3), ptr, ptr, 16, -1, None, 0, None) == 1
    assert L.rdetr_lsap_solve(ptr, one(3), one(3), ptr, ptr, 16, 0, None, 0, None) == 0
    assert L.rdetr_lsap_solve(ptr, one(-3), one(3), ptr, ptr, 16, 1, None, 0, None) == 1
    assert L.rdetr_lsap_solve(None, one(3), one(3), ptr, ptr, 16, 1, None, 0, None) == 1 and b"null" in L.rdetr_last_error()
    assert L.rdetr_match_cost(ptr, ptr, ptr, ptr, ptr, one(3), one(3), 0, 1.0, 1.0, 1.0, 0.25, 2.0, 1, None) == 1
    assert L.rdetr_match_cost(ptr, ptr, ptr, ptr, ptr, one(0), one(3), 91, 1.0, 1.0, 1.0, 0.25, 2.0, 1, None) == 0   # nothing to do
    assert L.rdetr_match_cost(ptr, ptr, ptr, ptr, ptr, one(1 << 20), one(1 << 20), 91, 1.0, 1.0, 1.0, 0.25, 2.0, 1, None) == 1
    with pytest.raises(_lib.RdetrOpsError):
        _lib.check(2, "x")


def test_no_cpu_fallback():
    inp = workloads.make_msda_inputs(workloads.MSDA_SHAPES["msda_tiny"], "U")
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.ms_deform_attn(inp["value"], inp["spatial_shapes"], inp["level_start_index"],
                           inp["sampling_locations"], inp["attention_weights"])
    r = workloads.make_rel_inputs(workloads.REL_SHAPES["rel_tiny"])
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.position_relation_bias(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"])
    m = modules.MultiScaleDeformableAttention()
    with pytest.raises((NotImplementedError, RuntimeError)):
        m(torch.zeros(1, 3, 256), torch.zeros(1, 3, 4, 2), torch.zeros(1, 126 + 24 + 6 + 2, 256),
          torch.tensor([[9, 14], [4, 6], [2, 3], [1, 2]]), torch.tensor([0, 126, 150, 156]), None)


def test_fake_kernels_give_shapes_and_dtypes():
    B, S, M, D, L, Nq, P = 2, 50, 8, 32, 3, 7, 4
    for dt in (torch.float32, torch.bfloat16):
        v = torch.empty(B, S, M, D, dtype=dt, device="meta")
        ss = torch.empty(L, 2, dtype=torch.int64, device="meta")
        lsi = torch.empty(L, dtype=torch.int64, device="meta")
        loc = torch.empty(B, Nq, M, L, P, 2, device="meta")
        attn = torch.empty(B, Nq, M, L, P, device="meta")
        out = torch.ops.rdetr.msda_forward(v, ss, lsi, loc, attn)
        assert out.shape == (B, Nq, M * D) and out.dtype == dt
        gv, gl, ga = torch.ops.rdetr.msda_backward(v, ss, lsi, loc, attn, out)
        assert gv.shape == v.shape and gv.dtype == dt and gl.shape == loc.shape and ga.shape == attn.shape
        ref = torch.empty(B, Nq, L, 4, device="meta")
        off = torch.empty(B, Nq, M, L, P, 2, dtype=dt, device="meta")
        z = torch.empty(B, Nq, M, L * P, dtype=dt, device="meta")
        out = torch.ops.rdetr.msda_fused_forward(v, ss, lsi, ref, off, z, None)
        assert out.shape == (B, Nq, M * D) and out.dtype == dt
        gv, go, gz = torch.ops.rdetr.msda_fused_backward(v, ss, lsi, ref, off, z, None, out)
        assert gv.shape == v.shape and go.shape == off.shape and go.dtype == dt and gz.shape == z.shape
    src = torch.empty(2, 37, 4, device="meta")
    tgt = torch.empty(2, 70, 4, device="meta")
    w, b, d = torch.empty(8, 64, device="meta"), torch.empty(8, device="meta"), torch.empty(8, device="meta")
    out, bits = torch.ops.rdetr.relation_forward(src, tgt, w, b, d, 100.0, 1e-5, None, False)
    assert out.shape == (2, 8, 37, 70) and bits.shape == (2, 37, 3, 8) and bits.dtype == torch.int32
    gw, gb = torch.ops.rdetr.relation_backward(src, tgt, d, 100.0, 1e-5, out, bits, 8, False)
    assert gw.shape == (8, 64) and gb.shape == (8,)


def test_function_apply_signature_and_step_check():
    v = torch.zeros(3, 4, 8, 32)
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        ops.MultiScaleDeformableAttnFunction.apply(v, None, None, None, None, 2)


MSDA_KEYS = ["attention_weights.bias", "attention_weights.weight", "output_proj.bias", "output_proj.weight",
             "sampling_offsets.bias", "sampling_offsets.weight", "value_proj.bias", "value_proj.weight"]


def test_module_parameter_names_and_shapes():
    m = rd.MultiScaleDeformableAttention(256, 4, 8, 4)
    sd = m.state_dict()
    assert sorted(sd) == MSDA_KEYS
    assert sd["sampling_offsets.weight"].shape == (8 * 4 * 4 * 2, 256) and sd["attention_weights.weight"].shape == (128, 256)
    assert torch.all(sd["sampling_offsets.weight"] == 0) and torch.all(sd["attention_weights.bias"] == 0)
    # offset bias = per-head direction * (point index + 1)   (ms_deform_attn.py:269-278)
    bias = sd["sampling_offsets.bias"].view(8, 4, 4, 2)
    assert torch.allclose(bias, workloads.grid_init(8, 4, 4))
    r = rd.PositionRelationEmbedding(16, 8)
    assert sorted(r.state_dict()) == ["pos_proj.0.bias", "pos_proj.0.weight"]
    assert r.pos_proj[0].weight.shape == (8, 64, 1, 1)
    with pytest.raises(NotImplementedError):
        rd.PositionRelationEmbedding(16, 8, activation_layer=torch.nn.GELU)
    with pytest.raises(ValueError):
        rd.MultiScaleDeformableAttention(250, 4, 8, 4)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted (GPU box)")
def test_state_dicts_interchange_with_the_reference_modules():
    _, PRE, _, MSDA = ref_import.load()
    torch.manual_seed(0)
    ref = MSDA(256, 5, 8, 4)
    torch.manual_seed(0)
    ours = rd.MultiScaleDeformableAttention(256, 5, 8, 4)
    assert list(ref.state_dict()) == list(ours.state_dict())
    for k, v in ref.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k  # same init under the same seed
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    rp, op = PRE(16, 8), rd.PositionRelationEmbedding(16, 8)
    op.load_state_dict(rp.state_dict(), strict=True)
    rp.load_state_dict(op.state_dict(), strict=True)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted (GPU box)")
def test_install_rebinds_reference_names():
    from relation_detr_b200 import install
    ref_import.load()
    import models.bricks.ms_deform_attn as ref_msda
    import models.bricks.relation_transformer as ref_rt
    import models.matcher.hungarian_matcher as ref_matcher
    saved = (ref_msda.MultiScaleDeformableAttention, ref_msda.MultiScaleDeformableAttnFunction,
             ref_rt.MultiScaleDeformableAttention, ref_rt.PositionRelationEmbedding)
    saved_matcher = ref_matcher.HungarianMatcher
    try:
        assert "models.matcher.hungarian_matcher.HungarianMatcher" not in install.install(matcher_too=False)
        assert ref_matcher.HungarianMatcher is saved_matcher
        rebound = install.install()
        assert "models.bricks.relation_transformer.PositionRelationEmbedding" in rebound
        assert "models.matcher.hungarian_matcher.HungarianMatcher" in rebound
        assert ref_matcher.HungarianMatcher is rd.HungarianMatcher
        import inspect
        want = inspect.signature(saved_matcher.__init__).parameters
        have = inspect.signature(rd.HungarianMatcher.__init__).parameters
        assert list(want) == list(have)[:len(want)] and all(want[k].default == have[k].default for k in want)
        assert list(inspect.signature(saved_matcher.forward).parameters) == list(inspect.signature(rd.HungarianMatcher.forward).parameters)
        assert ref_rt.MultiScaleDeformableAttention is rd.MultiScaleDeformableAttention
        layer = ref_rt.RelationTransformerEncoderLayer(256, 1024, 0.0, 8, torch.nn.ReLU(), 4, 4) \
            if hasattr(ref_rt, "RelationTransformerEncoderLayer") else None
        if layer is not None:
            assert isinstance(layer.self_attn, rd.MultiScaleDeformableAttention)
    finally:
        (ref_msda.MultiScaleDeformableAttention, ref_msda.MultiScaleDeformableAttnFunction,
         ref_rt.MultiScaleDeformableAttention, ref_rt.PositionRelationEmbedding) = saved
        ref_matcher.HungarianMatcher = saved_matcher


def test_shard_range_tiles_exactly():
    for total in (0, 1, 7, 8, 13, 64):
        for world in (1, 2, 3, 8):
            spans = [dist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dist.shard_range(4, 2, 2)


def test_workload_byte_counts_match_baseline_md():
    s = workloads.MSDA_SHAPES["msda_enc_800x1333_b8"]
    fwd, bwd = s.algorithmic_bytes(4)
    assert (s.S, s.Nq) == (22323, 22323)
    assert round(fwd / 1e6, 1) == 640.0 and round(bwd / 1e6, 1) == 1097.2
    f16, b16 = s.algorithmic_bytes(2)
    assert round(f16 / 1e6, 1) == 457.2 and round(b16 / 1e6, 1) == 822.9
    assert workloads.MSDA_SHAPES["msda_enc_1200x2000_b1"].S == 204098
    assert round(workloads.REL_SHAPES["rel_900_b8"].algorithmic_bytes()[0] / 1e6, 1) == 207.4


def test_generators_are_deterministic_on_cpu():
    a = workloads.make_msda_inputs(workloads.MSDA_SHAPES["msda_tiny"], "D", seed=3)
    b = workloads.make_msda_inputs(workloads.MSDA_SHAPES["msda_tiny"], "D", seed=3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    r1 = workloads.make_rel_inputs(workloads.REL_SHAPES["rel_tiny"], seed=5)
    r2 = workloads.make_rel_inputs(workloads.REL_SHAPES["rel_tiny"], seed=5)
    assert all(torch.equal(r1[k], r2[k]) for k in r1)
    strict = workloads.make_loc(workloads.MSDA_SHAPES["msda_tiny"], "strict", seed=1)
    wh = torch.tensor([[w, h] for h, w in workloads.MSDA_SHAPES["msda_tiny"].levels], dtype=torch.float32)
    px = strict * wh[None, None, None, :, None, :] - 0.5
    frac = px - px.floor()
    assert ((frac > 0.004) & (frac < 0.996)).all()  # the strict generator keeps samples off pixel boundaries


def test_graph_capture_helpers_host_logic():
    """relation_detr_b200.graphs on the host: the positional wrapper forwards keyword arguments to the CLASS's forward (the
    instance attribute is the dispatcher), the signature check tells calls apart, release() restores the instance."""
    import torch
    from torch import nn

    from relation_detr_b200 import graphs

    class Inner(nn.Module):
        def forward(self, a, b=None, flag=False):
            return a * 2 + (0 if b is None else b) + (1 if flag else 0)

    inner = Inner()
    inner.forward = lambda *a, **k: (_ for _ in ()).throw(AssertionError("the wrapper must bypass the instance attribute"))
    w = graphs._Positional(inner, ["a", "b"], {"flag": True})
    x, y = torch.ones(3), torch.full((3,), 5.0)
    assert torch.equal(w(x, y), x * 2 + y + 1)
    sig = graphs._Signature(["a", "b"], {"flag": True}, [x, y])
    assert sig.matches({"a": x, "b": y, "flag": True})
    assert not sig.matches({"a": x, "b": y, "flag": False})
    assert not sig.matches({"a": x, "flag": True})
    assert not sig.matches({"a": torch.ones(4), "b": y, "flag": True})
    assert graphs.autocast_kwargs(torch.bfloat16)["cache_enabled"] is False
    h = graphs.GraphHandle()
    h._undo.append(lambda: inner.__dict__.pop("forward"))
    h.release()
    assert "forward" not in inner.__dict__ and torch.equal(inner(x), x * 2)


def test_install_as_first_importer_of_the_reference_is_undone_by_uninstall():
    """install() in a fresh interpreter, where it is the first to import the reference's modules: relation_transformer
    binds MultiScaleDeformableAttention from ms_deform_attn at import time, so every user must be imported before the
    first name is rebound, or uninstall() restores this package's class as "the original" (order-dependent failure of
    the GPU integration tests)."""
    from baseline import refmodel
    if not refmodel.available():
        pytest.skip("baseline/_ref not installed")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from baseline import refmodel\n"
        "from relation_detr_b200 import install, modules\n"
        "refmodel.activate()\n"
        "assert 'models.bricks.relation_transformer' not in sys.modules\n"
        "install.install()\n"
        "rt = sys.modules['models.bricks.relation_transformer']\n"
        "assert rt.MultiScaleDeformableAttention is modules.MultiScaleDeformableAttention\n"
        "install.uninstall()\n"
        "assert rt.MultiScaleDeformableAttention.__module__ == 'models.bricks.ms_deform_attn', rt.MultiScaleDeformableAttention\n"
        "assert rt.PositionRelationEmbedding.__module__ == 'models.bricks.relation_transformer'\n"
        "assert not install._saved\n"
        "print('ok')\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import subprocess
    import sys
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
