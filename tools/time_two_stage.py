"""Two-stage query selection: rdetr_two_stage_select (3 launches) and rdetr_topk_rows vs the reference's expression
(relation_transformer.py:90-96: sigmoid over all boxes, max, torch.topk, two gathers), forward and forward+backward.
python tools/time_two_stage.py -> one JSON line per case."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relation_detr_b200 import ops  # noqa: E402

DEV = "cuda:0"


def timed(fn, warm=5, iters=30):
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
    for B, S, k in ((2, 22323, 900), (2, 22323, 1500), (8, 22323, 900), (1, 204098, 900), (1, 204098, 1500)):
        g = torch.Generator(device=DEV).manual_seed(0)
        cls = (torch.randn((B, S, 91), device=DEV, generator=g) - 4.6).requires_grad_(True)
        box = torch.randn((B, S, 4), device=DEV, generator=g).requires_grad_(True)
        gc = torch.randn((B, k, 91), device=DEV, generator=g)
        gb = torch.randn((B, k, 4), device=DEV, generator=g)
        scores = cls.detach().max(-1)[0]

        def ref():
            coord = box.sigmoid()
            idx = torch.topk(cls.max(-1)[0], k, dim=1)[1].unsqueeze(-1)
            return cls.gather(1, idx.expand(-1, -1, 91)), coord.gather(1, idx.expand(-1, -1, 4))

        def ours():
            return ops.two_stage_select(cls, box, k)[:2]

        def fb(f):
            def run():
                cls.grad = box.grad = None
                a, b = f()
                ((a * gc).sum() + (b * gb).sum()).backward()
            return run

        with torch.no_grad():
            row = {"B": B, "S": S, "k": k, "ours_fwd_ms": round(timed(ours), 4), "reference_fwd_ms": round(timed(ref), 4),
                   "ours_topk_only_ms": round(timed(lambda: ops.topk_rows(scores, k)), 4),
                   "torch_topk_only_ms": round(timed(lambda: torch.topk(scores, k, dim=1)), 4)}
        row["ours_fwd_bwd_ms"] = round(timed(fb(ours)), 4)
        row["reference_fwd_bwd_ms"] = round(timed(fb(ref)), 4)
        row["class_logits_MB"] = round(B * S * 91 * 4 / 1e6, 1)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
