"""CPU restatement (numpy) of the reference's two-stage query selection -- the checker for ``csrc/topk.cu``.

TEST INFRASTRUCTURE ONLY (imported by ``tests/`` and ``oracle/make_golden_two_stage.py``; never by the product path).
Follows ``RelationTransformer.forward`` (upstream ``models/bricks/relation_transformer.py:88-96`` and ``:104-111``):

    enc_outputs_coord = (bbox_head(output_memory) + output_proposals).sigmoid()                   :89-90
    topk_index = torch.topk(enc_outputs_class.max(-1)[0], topk, dim=1)[1].unsqueeze(-1)            :94
    enc_outputs_class = enc_outputs_class.gather(1, topk_index.expand(-1, -1, num_classes))        :95
    enc_outputs_coord = enc_outputs_coord.gather(1, topk_index.expand(-1, -1, 4))                  :96

``torch.topk`` is ATen (third-party, torch 2.11.0 in this image): scores descending, NaN ranked above every number; the
order of EQUAL scores is unspecified there -- this restatement (and the kernel) returns them by ascending index.
PINNED against the reference itself: ``tests/golden/twostage_*.npz`` hold what ``RelationTransformer.forward`` returned
for its main and hybrid selections on seeded inputs (``oracle/make_golden_two_stage.py``), and
``tests/test_two_stage_oracle.py`` compares this file with them and with ``torch.topk`` on tie-free scores.
"""
from __future__ import annotations

import numpy as np


def topk_rows(scores: np.ndarray, k: int) -> np.ndarray:
    """indices [B, k] of the k largest entries of every row of ``scores [B, S]`` (descending, NaN first, ties by index)."""
    B, S = scores.shape
    if not 0 < k <= S:
        raise ValueError(f"selected index k out of range (k = {k}, {S} elements per row)")
    out = np.empty((B, k), np.int64)
    for b in range(B):
        s = scores[b]
        nan = np.isnan(s)
        # lexsort: last key is the primary one -- NaN first, then score descending, then index ascending
        order = np.lexsort((np.arange(S), -np.where(nan, 0.0, s), ~nan))
        out[b] = order[:k]
    return out


def sigmoid(x: np.ndarray) -> np.ndarray:
    return 1.0 / (1.0 + np.exp(-x))


def two_stage_select(class_logits: np.ndarray, coord_unact: np.ndarray, k: int, apply_sigmoid: bool = True):
    """-> (topk_class [B, k, C], topk_coord [B, k, 4], topk_index [B, k])"""
    with np.errstate(invalid="ignore"):
        idx = topk_rows(class_logits.max(-1), k)   # np.max propagates NaN like torch's max
    cls = np.take_along_axis(class_logits, idx[..., None], 1)
    box = np.take_along_axis(coord_unact, idx[..., None], 1)
    if apply_sigmoid:
        with np.errstate(over="ignore"):
            box = sigmoid(box)
    return cls, box, idx


def two_stage_select_backward(grad_class, grad_coord, topk_coord, topk_index, S: int, apply_sigmoid: bool = True):
    """Adjoint of the two gathers (and of the sigmoid on the selected boxes): -> (grad_class_logits, grad_coord_unact)."""
    B, K, C = grad_class.shape
    gc = np.zeros((B, S, C), grad_class.dtype)
    gb = np.zeros((B, S, 4), grad_coord.dtype)
    g = grad_coord * ((1.0 - topk_coord) * topk_coord) if apply_sigmoid else grad_coord
    for b in range(B):
        gc[b, topk_index[b]] = grad_class[b]
        gb[b, topk_index[b]] = g[b]
    return gc, gb
