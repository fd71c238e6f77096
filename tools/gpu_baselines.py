"""GPU-side baselines next to our kernels, same inputs, CUDA-event timed (not part of bench.py):

  (a) the reference's effective GPU path on this image: the grid_sample port (oracle/torch_port.py)
      + autograd, and the eager relation embedding;
  (b) the reference's own CUDA extension, unmodified, recompiled for sm_100a (oracle/_ref, built by
      oracle/build_ref_cuda.py), when present -- also used as a third parity check (fp32 and fp64).

    python tools/gpu_baselines.py > profiles/rNN_gpu_baselines.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import build_ref_cuda, torch_port  # noqa: E402
from relation_detr_b200 import ops, workloads  # noqa: E402

DEV = "cuda:0"


def timed(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    res = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__}
    refc = build_ref_cuda.load_prebuilt()
    res["reference_cuda_ext"] = "oracle/_ref present" if refc is not None else "absent"
    for name, kind in (("msda_enc_800x1333_b8", "S"), ("msda_enc_800x1333_b8", "U"), ("msda_dec_900_b8", "D")):
        shape = workloads.MSDA_SHAPES[name]
        inp = workloads.make_msda_inputs(shape, kind, seed=0, device=DEV)
        v, ss, lsi = inp["value"], inp["spatial_shapes"], inp["level_start_index"]
        loc, attn, go = inp["sampling_locations"], inp["attention_weights"], inp["grad_output"]
        fb, bb = shape.algorithmic_bytes(4)
        entry = {}

        def ours_fwd():
            return ops.msda_forward(v, ss, lsi, loc, attn)

        def ours_bwd():
            return ops.msda_backward(v, ss, lsi, loc, attn, go)

        entry["ours"] = {"fwd_ms": timed(ours_fwd), "bwd_ms": timed(ours_bwd)}

        def port_fwd():
            return torch_port.msda_grid_sample(v, ss, loc, attn)

        def port_fwd_bwd():
            vv, ll, aa = v.detach().requires_grad_(True), loc.detach().requires_grad_(True), attn.detach().requires_grad_(True)
            torch_port.msda_grid_sample(vv, ss, ll, aa).backward(go)

        f = timed(port_fwd)
        fbw = timed(port_fwd_bwd)
        entry["reference_grid_sample_path"] = {"fwd_ms": f, "bwd_ms": fbw - f}
        if refc is not None:
            entry["reference_cuda_kernel_sm100a"] = {
                "fwd_ms": timed(lambda: refc.ms_deform_attn_forward(v, ss, lsi, loc, attn, 64)),
                "bwd_ms": timed(lambda: refc.ms_deform_attn_backward(v, ss, lsi, loc, attn, go, 64)),
            }
            # parity of ours against the reference's own kernel (fp32) and its fp64 run
            out_r = refc.ms_deform_attn_forward(v, ss, lsi, loc, attn, 64)
            out_r64 = refc.ms_deform_attn_forward(v.double(), ss, lsi, loc.double(), attn.double(), 64)
            gv_r64, gl_r64, ga_r64 = refc.ms_deform_attn_backward(v.double(), ss, lsi, loc.double(), attn.double(), go.double(), 64)
            out_o = ours_fwd()
            gv_o, gl_o, ga_o = ours_bwd()
            rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
            bad = ((gl_o.double() - gl_r64).abs() > 1e-4 * gl_r64.abs().max()).double().mean()
            entry["parity_vs_reference_cuda_fp64"] = {
                "out_maxabs_ours": float((out_o.double() - out_r64).abs().max()),
                "out_maxabs_reference_fp32": float((out_r.double() - out_r64).abs().max()),
                "grad_value_rel": rel(gv_o, gv_r64), "grad_attn_rel": rel(ga_o, ga_r64),
                "grad_loc_frac_over_1e-4": float(bad),
            }
        for k in entry:
            if "fwd_ms" in entry[k]:
                t = entry[k]["fwd_ms"] + entry[k]["bwd_ms"]
                entry[k]["fwd_bwd_GBps"] = (fb + bb) / t / 1e6
        res[f"{name}_loc{kind}_f32"] = entry
        del inp, v, loc, attn, go
        torch.cuda.empty_cache()

    for name in ("rel_900_b8", "rel_1100_b8"):
        shape = workloads.REL_SHAPES[name]
        r = workloads.make_rel_inputs(shape, seed=0, device=DEV)
        dim_t = ops.relation_dim_t(16, 10000.0, DEV)
        entry = {}
        for fast in (False, True):
            def fwd():
                return ops.relation_forward(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"], dim_t, 100.0, 1e-5, None, fast)
            out, bits = fwd()

            def bwd():
                return ops.relation_backward(r["src_boxes"], r["tgt_boxes"], dim_t, 100.0, 1e-5, r["grad_output"], bits, 8, fast)
            entry["ours_fast" if fast else "ours_exact"] = {"fwd_ms": timed(fwd), "bwd_ms": timed(bwd)}

        def eager_fwd():
            return torch_port.rel_eager(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"])

        def eager_fwd_bwd():
            w = r["weight"].detach().requires_grad_(True)
            b = r["bias"].detach().requires_grad_(True)
            torch_port.rel_eager(r["src_boxes"], r["tgt_boxes"], w, b).backward(r["grad_output"])

        f = timed(eager_fwd)
        fbw = timed(eager_fwd_bwd)
        entry["reference_eager_path"] = {"fwd_ms": f, "bwd_ms": fbw - f, "note": "cudnn TF32 conv allowed (torch default)"}
        res[name] = entry
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
