"""Time one training step's worth of bipartite matching on the device solver
against the reference's way (cost on the GPU, ``.cpu()``, SciPy) -- hungarian_matcher.py:74-81.

A Relation-DETR step matches 14 prediction sets per image: 6 decoder layers + the encoder proposals with
900 queries, and the same again for the hybrid branch with 1500 queries against 6 copies of every target
(relation_detr.py:96-134, set_criterion.py:120-170).  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relation_detr_b200 as rd  # noqa: E402


def scipy_way(m, sets):
    from scipy.optimize import linear_sum_assignment
    out = []
    for pb, pl, gb, gl in sets:
        c = m.calculate_cost(pb, pl, gb, gl)
        r, c_ = linear_sum_assignment(c.cpu())          # the reference's line 80: copy + host solve
        out.append((torch.as_tensor(r), torch.as_tensor(c_)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--gt", type=int, nargs="*", default=[7, 15, 3, 30, 11, 1, 52, 9])
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    m = rd.HungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2)
    sets = []
    for b in range(a.batch):
        n = a.gt[b % len(a.gt)]
        gb = torch.cat([torch.rand(n, 2, generator=g) * 0.8 + 0.1, torch.rand(n, 2, generator=g) * 0.3 + 0.02], -1).to(dev)
        gl = torch.randint(0, 91, (n,), generator=g).to(dev)
        for nq, rep in ((900, 1), (1500, 6)):
            for _ in range(7):
                pb = torch.cat([torch.rand(nq, 2, generator=g) * 0.8 + 0.1, torch.rand(nq, 2, generator=g) * 0.3 + 0.02], -1).to(dev)
                pl = (torch.randn(nq, 91, generator=g) * 2 - 2).to(dev)
                sets.append((pb, pl, gb.repeat(rep, 1), gl.repeat(rep)))
    cols = list(zip(*sets))

    ref = scipy_way(m, sets)
    ours = m.match_batch(*cols)
    torch.cuda.synchronize()
    same = all(np.array_equal(r[0].numpy(), o[0].cpu().numpy()) and np.array_equal(r[1].numpy(), o[1].cpu().numpy())
               for r, o in zip(ref, ours))

    def wall(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / a.iters * 1e3

    t_ref = wall(lambda: scipy_way(m, sets))
    t_ours = wall(lambda: m.match_batch(*cols))
    m_eager = rd.HungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, fused_cost=False)
    t_ours_eager = wall(lambda: m_eager.match_batch(*cols))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        m.match_batch(*cols)
    e1.record()
    torch.cuda.synchronize()
    t_ours_dev = e0.elapsed_time(e1) / a.iters
    # solver alone, costs already computed
    costs = [m.calculate_cost(*s).float() for s in sets]
    rd.ops.lsap_solve(costs)
    torch.cuda.synchronize()
    samples = []
    for _ in range(a.iters):
        # a spin kernel keeps the GPU busy while the host enqueues the call: the events bracket device time only
        torch.cuda._sleep(30_000_000)
        e0.record()
        rd.ops.lsap_solve(costs)
        e1.record()
        torch.cuda.synchronize()
        samples.append(e0.elapsed_time(e1))
    t_solver = sorted(samples)[len(samples) // 2]
    t0 = time.perf_counter()
    host = [c.cpu().numpy() for c in costs]
    from scipy.optimize import linear_sum_assignment
    t_copy = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    for c in host:
        linear_sum_assignment(c)
    t_scipy = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"workload": f"matching of one step, batch {a.batch}: {len(sets)} problems (900 x G and 1500 x 6G)",
                      "gt_per_image": [a.gt[b % len(a.gt)] for b in range(a.batch)], "identical_to_scipy": bool(same),
                      "reference_way_ms": round(t_ref, 3), "device_matcher_ms": round(t_ours, 3),
                      "device_matcher_gpu_time_ms": round(t_ours_dev, 3), "device_matcher_eager_cost_ms": round(t_ours_eager, 3),
                      "device_solver_only_ms": round(t_solver, 3), "scipy_solve_only_ms": round(t_scipy, 3),
                      "d2h_copies_ms": round(t_copy, 3), "host_syncs_removed_per_step": len(sets)}))


if __name__ == "__main__":
    main()
