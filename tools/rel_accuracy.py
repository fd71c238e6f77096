"""Error of the relation kernels (EXACT, FAST) and of the reference's own fp32 evaluation against the fp64
truth (C oracle), B=1, N=900 -- decides which mode the module may default to."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle, torch_port  # noqa: E402
from relation_detr_b200 import ops, workloads  # noqa: E402

dev = "cuda:0"
res = {}
for seed in (0, 1):
    r = workloads.make_rel_inputs(workloads.RelShape("t", 1, 900, 900), seed=seed)
    dim_t = torch_port.relation_dim_t().numpy().astype(np.float64)
    a64 = [r[k].numpy().astype(np.float64) for k in ("src_boxes", "tgt_boxes", "weight", "bias")]
    truth = c_oracle.rel_forward(*a64, dim_t)
    gw64, gb64 = c_oracle.rel_backward(*a64, dim_t, r["grad_output"].numpy().astype(np.float64))
    d = {k: v.to(dev) for k, v in r.items()}
    torch.backends.cudnn.allow_tf32 = False
    eager32 = torch_port.rel_eager(d["src_boxes"], d["tgt_boxes"], d["weight"], d["bias"]).cpu().numpy()
    torch.backends.cudnn.allow_tf32 = True
    eager_tf32 = torch_port.rel_eager(d["src_boxes"], d["tgt_boxes"], d["weight"], d["bias"]).cpu().numpy()
    entry = {"reference_eager_fp32": {"max": float(np.abs(eager32 - truth).max()), "mean": float(np.abs(eager32 - truth).mean())},
             "reference_eager_tf32_default": {"max": float(np.abs(eager_tf32 - truth).max()), "mean": float(np.abs(eager_tf32 - truth).mean())}}
    for fast in (False, True):
        w = d["weight"].clone().requires_grad_(True)
        b = d["bias"].clone().requires_grad_(True)
        out = ops.position_relation_bias(d["src_boxes"], d["tgt_boxes"], w, b, fast=fast)
        out.backward(d["grad_output"])
        o = out.detach().cpu().numpy()
        entry["ours_fast" if fast else "ours_exact"] = {
            "max": float(np.abs(o - truth).max()), "mean": float(np.abs(o - truth).mean()),
            "grad_weight_rel": float(np.abs(w.grad.cpu().numpy() - gw64).max() / np.abs(gw64).max()),
            "grad_bias_rel": float(np.abs(b.grad.cpu().numpy() - gb64).max() / np.abs(gb64).max())}
    res[f"seed{seed}"] = entry
print(json.dumps(res, indent=1))
