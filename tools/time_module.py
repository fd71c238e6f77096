"""Times the MSDA *module* (4 Linears + prologue + kernel) fwd+bwd at the encoder shape, fused vs unfused
prologue, fp32 and bf16 autocast; and the reference's effective path (oracle-backed module) for context."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import relation_detr_b200 as rd  # noqa: E402
from relation_detr_b200 import workloads  # noqa: E402

dev = "cuda:0"
res = {}
for name, nq_kind in (("msda_enc_800x1333_b8", "enc"), ("msda_dec_900_b8", "dec")):
    shape = workloads.MSDA_SHAPES[name]
    torch.manual_seed(0)
    mod = rd.MultiScaleDeformableAttention(256, 4, 8, 4).to(dev)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.01)
    ss, lsi = workloads.shape_tensors(shape.levels, dev)
    mem = torch.randn((shape.batch, shape.S, 256), device=dev, requires_grad=True)
    mask = torch.zeros((shape.batch, shape.S), dtype=torch.bool, device=dev)
    if nq_kind == "enc":
        query = torch.randn((shape.batch, shape.S, 256), device=dev, requires_grad=True)
        ref = workloads.full_reference_points(shape.levels, dev)[None, :, None, :].expand(shape.batch, -1, 4, -1).contiguous()
    else:
        query = torch.randn((shape.batch, shape.Nq, 256), device=dev, requires_grad=True)
        ref = workloads.make_boxes(shape.batch, shape.Nq, 0, dev)[:, :, None, :].expand(-1, -1, 4, -1).contiguous()

    def step():
        out = mod(query, ref, mem, ss, lsi, mask)
        out.sum().backward()

    def timed(iters=10):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / iters, 4)

    entry = {}
    for fused in (True, False):
        mod.fused_prologue = fused
        key = "fused" if fused else "unfused"
        entry[f"{key}_fp32_ms"] = timed()
        torch.backends.cuda.matmul.allow_tf32 = True
        entry[f"{key}_tf32_matmul_ms"] = timed()
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.autocast("cuda", dtype=torch.bfloat16):
            entry[f"{key}_bf16_autocast_ms"] = timed()
    res[name + "_module_fwd_bwd"] = entry
print(json.dumps(res, indent=1))
