"""memory_fusion input Linear: the K-split tcgen05 GEMM vs the chain upstream runs (torch.cat -> Linear -> ReLU), forward and
forward+backward, at the encoder's sizes.  python tools/time_memory_fusion.py [--profile] -> one JSON line per case."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relation_detr_b200 import ops  # noqa: E402

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
    profile = "--profile" in sys.argv
    for B, S in ((8, 22323), (2, 22323)):
        g = torch.Generator(device=DEV).manual_seed(0)
        srcs = [torch.randn((B, S, 256), device=DEV, generator=g).requires_grad_(True) for _ in range(7)]
        lin = torch.nn.Linear(1792, 256).to(DEV)
        go = torch.randn((B, S, 256), device=DEV, generator=g)

        def ours():
            return ops.memory_fusion_linear(srcs, lin.weight, lin.bias, True)

        def ours_autocast():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return ops.memory_fusion_linear(srcs, lin.weight, lin.bias, True)

        def ref_fp32():
            return torch.relu(lin(torch.cat(srcs, -1)))

        def ref_bf16():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return torch.relu(lin(torch.cat(srcs, -1)))

        def ref_tf32():
            torch.backends.cuda.matmul.allow_tf32 = True
            try:
                return torch.relu(lin(torch.cat(srcs, -1)))
            finally:
                torch.backends.cuda.matmul.allow_tf32 = False

        def fb(f):
            def run():
                for t in srcs:
                    t.grad = None
                lin.zero_grad()
                f().backward(go.to(f().dtype) if False else go)
            return run

        if profile:
            ours()
            torch.cuda.synchronize()
            print("profiled")
            return
        M = B * S
        flops = 2.0 * M * 1792 * 256
        bytes_alg = (7 * M * 256 + M * 256 + 1792 * 256) * 4
        res = {"B": B, "S": S, "M": M, "gflop": round(flops / 1e9, 1), "algorithmic_MB": round(bytes_alg / 1e6, 1)}
        for name, f in (("ours_tcgen05_tf32", ours), ("ours_tcgen05_tf32_under_autocast", ours_autocast), ("torch_cat_linear_fp32", ref_fp32), ("torch_cat_linear_tf32", ref_tf32),
                        ("torch_cat_linear_bf16_autocast", ref_bf16)):
            torch.cuda.reset_peak_memory_stats()
            t = timed(f)
            entry = {"fwd_ms": round(t, 4), "TFLOPs": round(flops / t / 1e9, 1), "GBps_algorithmic": round(bytes_alg / t / 1e6, 1)}
            try:
                entry["fwd_bwd_ms"] = round(timed(fb(f), 2, 5), 4)
            except Exception as e:  # noqa: BLE001
                entry["fwd_bwd_error"] = str(e)[:120]
            entry["peak_mem_MB"] = round(torch.cuda.max_memory_allocated() / 2**20, 1)
            res[name] = entry
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
