"""Fused relation attention vs the unfused chain it replaces (rel kernel -> masked_fill -> SDPA with a float mask), fp32,
forward and forward+backward, at the decoder's sizes.  python tools/time_rel_attention.py -> one JSON line per size."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relation_detr_b200 import ops, workloads  # noqa: E402

DEV = "cuda:0"


def timed(fn, warm=3, iters=10):
    for _ in range(warm):
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
    for B, N, dn in ((8, 900, 0), (8, 1100, 200), (2, 1100, 200), (1, 2900, 2000)):   # (2, 1100) = the training shape per GPU
        g = torch.Generator(device=DEV).manual_seed(0)
        q, k, v = (torch.randn((B, 8, N, 32), device=DEV, generator=g).requires_grad_(True) for _ in range(3))
        src, tgt = workloads.make_boxes(B, N, 1, DEV), workloads.make_boxes(B, N, 2, DEV)
        w, b = workloads.make_rel_params(8, 64, 0, DEV)
        w.requires_grad_(True)
        b.requires_grad_(True)
        mask = workloads.cdn_attn_mask(N - dn, 10, dn // 10, DEV) if dn else None
        go = torch.randn((B, 8, N, 32), device=DEV, generator=g)

        def fused_fwd():
            return ops.relation_attention(q, k, v, src, tgt, w, b, attn_mask=mask)

        def unfused_fwd():
            bias = ops.position_relation_bias(src, tgt, w, b, attn_mask=mask, fast=True)
            return torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=bias)

        def unfused_ref_fwd():  # what the reference's decoder does: separate masked_fill_ on the materialised bias
            bias = ops.position_relation_bias(src, tgt, w, b, fast=True).flatten(0, 1)
            if mask is not None:
                bias.masked_fill_(mask, float("-inf"))
            return torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=bias.view(B, 8, N, N))

        def fb(f):
            def run():
                for t in (q, k, v, w, b):
                    t.grad = None
                f().backward(go)
            return run

        res = {"B": B, "N": N, "masked": bool(dn)}
        for name, f in (("fused", fused_fwd), ("unfused_fused_mask", unfused_fwd), ("unfused_as_reference", unfused_ref_fwd)):
            torch.cuda.reset_peak_memory_stats()
            res[name] = {"fwd_ms": round(timed(f), 4), "fwd_bwd_ms": round(timed(fb(f)), 4),
                         "peak_mem_MB": round(torch.cuda.max_memory_allocated() / 2**20, 1)}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
