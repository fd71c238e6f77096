"""Generate tests/golden/matcher_*.npz from the REFERENCE's HungarianMatcher (build container only).

TEST INFRASTRUCTURE ONLY.  Run as ``python -m oracle.make_golden_matcher`` where /root/reference is
mounted.  Each file holds seeded predictions / ground truth, the cost matrix the reference computes
(models/matcher/hungarian_matcher.py:62-72, CPU fp32) and the index pair its ``forward`` returns
(:74-91, through scipy.optimize.linear_sum_assignment).  They pin oracle/lsap_oracle.c on the CPU and
the device solver (rdetr_lsap_solve) plus relation-detr_b200/matcher.py on the GPU box.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name, queries, ground-truth boxes, classes, hybrid repeat of the targets (relation_detr.py:131-132),
# mixed_match, gt_copy, seed
CASES = [
    ("matcher_single", 300, 17, 91, 1, False, 1, 0),
    ("matcher_crowd", 900, 93, 91, 1, False, 1, 1),
    ("matcher_hybrid6", 500, 11, 91, 6, False, 1, 2),     # six identical copies of every target: exact ties
    ("matcher_more_gt", 20, 37, 91, 1, False, 1, 3),      # more boxes than queries: no transposition
    ("matcher_no_gt", 50, 0, 91, 1, False, 1, 4),
    ("matcher_mixed", 200, 9, 91, 1, True, 4, 5),          # Align-DETR mixed assignment
    ("matcher_dup_queries", 64, 8, 20, 1, False, 1, 6),    # repeated predictions: ties between queries
]


def make_case(nq, ng, ncls, repeat, seed, dup_queries=False):
    g = torch.Generator().manual_seed(seed)
    pred_boxes = torch.cat([torch.rand(nq, 2, generator=g) * 0.8 + 0.1, torch.rand(nq, 2, generator=g) * 0.3 + 0.02], -1)
    pred_logits = torch.randn(nq, ncls, generator=g) * 2 - 2
    gt_boxes = torch.cat([torch.rand(ng, 2, generator=g) * 0.8 + 0.1, torch.rand(ng, 2, generator=g) * 0.3 + 0.02], -1)
    gt_labels = torch.randint(0, ncls, (ng,), generator=g)
    if dup_queries:
        pred_boxes[nq // 2:] = pred_boxes[: nq - nq // 2]
        pred_logits[nq // 2:] = pred_logits[: nq - nq // 2]
    if repeat > 1:
        gt_boxes, gt_labels = gt_boxes.repeat(repeat, 1), gt_labels.repeat(repeat)
    return pred_boxes, pred_logits, gt_boxes, gt_labels


def main():
    assert ref_import.available(), "reference tree not mounted"
    if ref_import.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_import.REFERENCE_ROOT)
    from models.matcher.hungarian_matcher import HungarianMatcher

    for name, nq, ng, ncls, repeat, mixed, gt_copy, seed in CASES:
        pb, pl, gb, gl = make_case(nq, ng, ncls, repeat, seed, dup_queries="dup" in name)
        m = HungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, mixed_match=mixed)   # configs/relation_detr/*.py:75-81
        cost = m.calculate_cost(pb, pl, gb, gl)
        src, tgt = m(pb, pl, gb, gl, gt_copy=gt_copy)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), pred_boxes=pb.numpy(), pred_logits=pl.numpy(),
                            gt_boxes=gb.numpy(), gt_labels=gl.numpy(), cost=cost.numpy(), src_ind=src.numpy().astype(np.int64),
                            tgt_ind=tgt.numpy().astype(np.int64), mixed=np.array(mixed), gt_copy=np.array(gt_copy))
        print(f"{name}: cost {tuple(cost.shape)} -> {len(src)} pairs")


if __name__ == "__main__":
    main()
