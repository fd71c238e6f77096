"""Smallest inputs that exercise every kernel (ragged tails, out-of-range samples, masks) for
compute-sanitizer (one tool per gpurun call):  compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relation_detr_b200 import ops, workloads  # noqa: E402

dev = "cuda:0"
for L, P, M, Nq, B in ((3, 4, 8, 37, 2), (5, 4, 8, 5, 1), (1, 1, 3, 1, 1)):
    levels = tuple((max(1, 9 >> i), max(2, 13 >> i)) for i in range(L))
    shape = workloads.MsdaShape("s", B, levels, Nq, heads=M, points=P)
    inp = workloads.make_msda_inputs(shape, "oob", seed=L, device=dev)
    for dt in (torch.float32, torch.bfloat16):
        v = inp["value"].to(dt)
        out = ops.msda_forward(v, inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
        ops.msda_backward(v, inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"],
                          inp["grad_output"].to(dt))
for B, N1, N2 in ((2, 37, 29), (1, 1, 1), (1, 65, 33), (1, 130, 70)):
    r = workloads.make_rel_inputs(workloads.RelShape("s", B, N1, N2), seed=N1, device=dev)
    dim_t = ops.relation_dim_t(16, 10000.0, dev)
    mask = torch.rand((N1, N2), device=dev) > 0.5
    for fast in (False, True):
        out, bits = ops.relation_forward(r["src_boxes"], r["tgt_boxes"], r["weight"], r["bias"], dim_t, 100.0, 1e-5, mask, fast)
        ops.relation_backward(r["src_boxes"], r["tgt_boxes"], dim_t, 100.0, 1e-5, r["grad_output"], bits, 8, fast)
torch.cuda.synchronize()
print("sanitize_case done")
