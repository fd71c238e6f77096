"""Short driver for ncu: a couple of launches of each hot-path kernel at the benchmark shapes.

    python tools/profile_ops.py [msda|rel|all] [--loc S|U] [--dtype f32|bf16] [--iters N]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relation_detr_b200 import ops, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", nargs="?", default="all")
ap.add_argument("--loc", default="S")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--shape", default="msda_enc_800x1333_b8")
ap.add_argument("--fused", action="store_true", help="run the fused-prologue kernels instead of the plain ones")
a = ap.parse_args()
dev = "cuda:0"
if a.what in ("msda", "all"):
    shape = workloads.MSDA_SHAPES[a.shape]
    inp = workloads.make_msda_inputs(shape, a.loc, seed=0, device=dev)
    dt = torch.float32 if a.dtype == "f32" else torch.bfloat16
    v, go = inp["value"].to(dt), inp["grad_output"].to(dt)
    if a.fused:
        g = torch.Generator(device=dev).manual_seed(0)
        B, S, Nq, M, L, P = shape.batch, shape.S, shape.Nq, shape.heads, shape.L, shape.points
        off = (workloads.grid_init(M, L, P).to(dev)[None, None] + torch.randn((B, Nq, M, L, P, 2), device=dev, generator=g)).to(dt)
        logits = torch.randn((B, Nq, M, L * P), device=dev, generator=g).to(dt)
        ref = workloads.full_reference_points(shape.levels, dev)[None, :, None, :].expand(B, -1, L, -1).contiguous()
        for _ in range(a.iters):
            ops.msda_fused_forward(v, inp["spatial_shapes"], inp["level_start_index"], ref, off, logits, None)
            ops.msda_fused_backward(v, inp["spatial_shapes"], inp["level_start_index"], ref, off, logits, None, go)
        torch.cuda.synchronize()
        print("profile_ops done")
        sys.exit(0)
    for _ in range(a.iters):
        out = ops.msda_forward(v, inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
        ops.msda_backward(v, inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"], go)
    torch.cuda.synchronize()
if a.what in ("rel", "all"):
    r = workloads.make_rel_inputs(workloads.REL_SHAPES["rel_900_b8"], seed=0, device=dev)
    dim_t = ops.relation_dim_t(16, 10000.0, dev)
    for fast in (False, True):
        for _ in range(a.iters):
            out, bits = ops.relation_forward(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"], dim_t, 100.0, 1e-5, None, fast)
            ops.relation_backward(r["src_boxes"], r["tgt_boxes"], dim_t, 100.0, 1e-5, r["grad_output"], bits, 8, fast)
    torch.cuda.synchronize()
print("profile_ops done")
