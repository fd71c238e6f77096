"""Parity report in the form SURVEY.md 8(c) asks for: per tensor, ours vs oracle(fp32), ours vs oracle(fp64) and
the oracle's own fp32-vs-fp64 noise floor, on the same device, at the benchmark shapes.

    python tools/parity_report.py > profiles/rNN_parity_report.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_port  # noqa: E402
from relation_detr_b200 import ops, workloads  # noqa: E402

DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False  # keep the eager relation oracle in fp32


def maxabs(a, b):
    return float((a.double() - b.double()).abs().max())


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))


def frac_over(a, b, tol=1e-4):
    d = (a.double() - b.double()).abs()
    return float((d > tol * b.double().abs().max()).double().mean())


def oracle_msda(inp, dtype):
    v = inp["value"].detach().to(dtype).clone().requires_grad_(True)
    loc = inp["sampling_locations"].detach().to(dtype).clone().requires_grad_(True)
    attn = inp["attention_weights"].detach().to(dtype).clone().requires_grad_(True)
    out = torch_port.msda_grid_sample(v, inp["spatial_shapes"], loc, attn)
    out.backward(inp["grad_output"].to(dtype))
    return dict(out=out.detach(), grad_value=v.grad, grad_loc=loc.grad, grad_attn=attn.grad)


def ours_msda(inp, dtype=torch.float32):
    v = inp["value"].to(dtype)
    out = ops.msda_forward(v, inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
    gv, gl, ga = ops.msda_backward(v, inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"],
                                   inp["attention_weights"], inp["grad_output"].to(dtype))
    return dict(out=out, grad_value=gv, grad_loc=gl, grad_attn=ga)


def main():
    rep = {"device": torch.cuda.get_device_name(0), "columns": "ours_vs_oracle32 / ours_vs_oracle64 / oracle32_vs_oracle64",
           "metrics": "out: max-abs; grads: max-abs / max-abs(ref); grad_loc additionally: fraction of elements off by > 1e-4 * max"}
    for name, kind in (("msda_enc_800x1333_b1", "U"), ("msda_enc_800x1333_b1", "S"), ("msda_enc_800x1333_b1", "strict"),
                       ("msda_dec_900_b8", "D"), ("msda_enc_1200x2000_b1", "S")):
        base = workloads.MSDA_SHAPES[name]
        shape = base if base.batch <= 2 else workloads.MsdaShape(base.name, 2, base.levels, base.num_query)
        if name == "msda_enc_1200x2000_b1":
            shape = workloads.MsdaShape(base.name, 1, base.levels, 20000)
        inp = workloads.make_msda_inputs(shape, kind, seed=0, device=DEV)
        o32, o64, mine = oracle_msda(inp, torch.float32), oracle_msda(inp, torch.float64), ours_msda(inp)
        e = {"out_maxabs": [maxabs(mine["out"], o32["out"]), maxabs(mine["out"], o64["out"]), maxabs(o32["out"], o64["out"])]}
        for k in ("grad_value", "grad_attn", "grad_loc"):
            e[k + "_rel"] = [relmax(mine[k], o32[k]), relmax(mine[k], o64[k]), relmax(o32[k], o64[k])]
        e["grad_loc_frac_over_1e-4"] = [frac_over(mine["grad_loc"], o32["grad_loc"]), frac_over(mine["grad_loc"], o64["grad_loc"]),
                                        frac_over(o32["grad_loc"], o64["grad_loc"])]
        # bf16 value path against the fp64 oracle fed the same bf16-rounded inputs
        inp16 = dict(inp, value=inp["value"].bfloat16().float(), grad_output=inp["grad_output"].bfloat16().float())
        r64 = oracle_msda(inp16, torch.float64)
        m16 = ours_msda(inp16, torch.bfloat16)
        e["bf16_vs_oracle64"] = {"out_rel": relmax(m16["out"], r64["out"]), "grad_value_rel": relmax(m16["grad_value"], r64["grad_value"]),
                                 "grad_attn_rel": relmax(m16["grad_attn"], r64["grad_attn"])}
        rep[f"{name}_loc{kind}_B{shape.batch}_Nq{shape.Nq}"] = e
        del inp, o32, o64, mine, inp16, r64, m16
        torch.cuda.empty_cache()

    for n in (900, 1100):
        r = workloads.make_rel_inputs(workloads.RelShape("t", 2, n, n), seed=0, device=DEV)
        dim_t = torch_port.relation_dim_t(device=DEV)
        res = {}

        def eager(dtype):
            w = r["weight"].to(dtype).clone().requires_grad_(True)
            b = r["bias"].to(dtype).clone().requires_grad_(True)
            out = torch_port.rel_eager(r["src_boxes"].to(dtype), r["tgt_boxes"].to(dtype), w, b, dim_t.to(dtype))
            out.backward(r["grad_output"].to(dtype))
            return out.detach(), w.grad, b.grad

        o32, o64 = eager(torch.float32), eager(torch.float64)
        for fast in (False, True):
            w = r["weight"].clone().requires_grad_(True)
            b = r["bias"].clone().requires_grad_(True)
            out = ops.position_relation_bias(r["src_boxes"], r["tgt_boxes"], w, b, fast=fast)
            out.backward(r["grad_output"])
            res["fast" if fast else "exact"] = {
                "out_maxabs": [maxabs(out, o32[0]), maxabs(out, o64[0]), maxabs(o32[0], o64[0])],
                "out_meanabs_vs_oracle64": float((out.double() - o64[0]).abs().mean()),
                "grad_weight_rel": [relmax(w.grad, o32[1]), relmax(w.grad, o64[1]), relmax(o32[1], o64[1])],
                "grad_bias_rel": [relmax(b.grad, o32[2]), relmax(b.grad, o64[2]), relmax(o32[2], o64[2])],
                "relu_sign_disagreements_vs_oracle64": int(((out > 0) != (o64[0] > 0)).sum())}
        rep[f"relation_B2_N{n}"] = res
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
