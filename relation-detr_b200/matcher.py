"""Device-resident ``HungarianMatcher`` -- the reference's matcher without its host round trip.

Mirrors ``models/matcher/hungarian_matcher.py`` of the reference (same constructor arguments, same
``calculate_*`` methods, same ``forward`` signature and result order).  Two things change.  The cost matrix
(upstream: ~30 eager kernels per problem, :40-72) comes from one ``rdetr_match_cost`` launch for all problems,
which returns the eager chain's values bit for bit (``fused_cost=False`` keeps the eager chain).  And line
80 / 87: instead of ``linear_sum_assignment(c.cpu())`` -- a device->host copy that stalls the stream once per
image and per decoder layer -- the matrix is handed to ``rdetr_lsap_solve`` and the index tensors stay on the
device.  Results are the pairs SciPy returns for the same matrix, ties included (tests/test_lsap_gpu.py).

Differences, deliberate:
* the two index tensors are CUDA int64 tensors (upstream: CPU tensors); every use in
  ``models/bricks/set_criterion.py`` (:53-54, :90-92, :110-117) indexes device tensors with them, which
  now needs no implicit host->device copy either;
* SciPy's ``ValueError`` for an infeasible matrix / a NaN entry cannot be raised without synchronising.
  ``check_status=True`` restores it (one sync per call); by default the solver's status word stays on the
  device in ``last_status`` and failed problems return ``-1`` indices;
* the fused cost kernel computes in float32 and up-casts bf16/fp16 predictions first; under autocast the reference
  evaluates the class term in the logits' reduced precision.  ``fused_cost=False`` keeps the eager chain and with it
  the reference's dtype behaviour (the solver takes the float32 image of whatever that chain produced, exactly);
* ``match_batch`` solves all images of one prediction set in a single launch (upstream: ``map`` over images).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import Tensor, nn

from . import ops


def _cxcywh_to_xyxy(b: Tensor) -> Tensor:
    # torchvision.ops.boxes._box_cxcywh_to_xyxy, operation for operation (same rounding)
    cx, cy, w, h = b.unbind(-1)
    return torch.stack((cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h), dim=-1)


def _generalized_box_iou(b1: Tensor, b2: Tensor) -> Tensor:
    # torchvision.ops.generalized_box_iou for floating-point boxes, operation for operation
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    union = area1[:, None] + area2 - inter
    iou = inter / union
    lti = torch.min(b1[:, None, :2], b2[:, :2])
    rbi = torch.max(b1[:, None, 2:], b2[:, 2:])
    whi = (rbi - lti).clamp(min=0)
    areai = whi[:, :, 0] * whi[:, :, 1]
    return iou - (areai - union) / areai


class HungarianMatcher(nn.Module):
    """Drop-in for ``models.matcher.hungarian_matcher.HungarianMatcher`` (hungarian_matcher.py:8-91)."""

    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, focal_alpha: float = 0.25,
                 focal_gamma: float = 2.0, mixed_match: bool = False, check_status: bool = False,
                 fused_cost: bool = True):
        super().__init__()
        self.cost_class = cost_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.focal_alpha = focal_alpha
        self.focal_gamma = focal_gamma
        self.mixed_match = mixed_match
        self.check_status = check_status
        self.fused_cost = fused_cost
        self.last_status = None
        self._flag_host = None
        self._flag_event = None
        self._flag_pending = False

    # -- cost terms: hungarian_matcher.py:40-72, same operations in the same order -----------------------
    def calculate_class_cost(self, pred_logits, gt_labels, **kwargs):
        # focal-style cost of assigning each query to each target's class; the evaluation order
        # ((scale * p^gamma) * log(. + 1e-6), positive minus negative term) fixes the float32 rounding
        p = pred_logits.sigmoid()
        gamma, alpha = self.focal_gamma, self.focal_alpha
        negative = -(1 - alpha) * p**gamma * (1 - p + 1e-6).log()
        positive = -alpha * (1 - p)**gamma * (p + 1e-6).log()
        return positive[:, gt_labels] - negative[:, gt_labels]

    def calculate_bbox_cost(self, pred_boxes, gt_boxes, **kwargs):
        return torch.cdist(pred_boxes, gt_boxes, p=1)      # L1 distance between cxcywh boxes

    def calculate_giou_cost(self, pred_boxes, gt_boxes, **kwargs):
        return -_generalized_box_iou(_cxcywh_to_xyxy(pred_boxes), _cxcywh_to_xyxy(gt_boxes))

    @torch.no_grad()
    def calculate_cost(self, pred_boxes: Tensor, pred_logits: Tensor, gt_boxes: Tensor, gt_labels: Tensor):
        terms = (self.cost_bbox * self.calculate_bbox_cost(pred_boxes, gt_boxes),
                 self.cost_class * self.calculate_class_cost(pred_logits, gt_labels),
                 self.cost_giou * self.calculate_giou_cost(pred_boxes, gt_boxes))
        return terms[0] + terms[1] + terms[2]              # bbox + class, then + giou: upstream's order (:71)

    # -- assignment ----------------------------------------------------------------------------------------
    def _prepare(self, c: Tensor, gt_copy: int) -> Tuple[Tensor, int]:
        if not self.mixed_match:
            return c.float(), 0
        gt_size = c.size(-1)
        num_queries = len(c)
        gt_copy = min(int(num_queries * 0.5 / gt_size), gt_copy) if gt_size > 0 else gt_copy
        return c.float().repeat(1, gt_copy), gt_size

    def _finish(self, src_ind: Tensor, tgt_ind: Tensor, gt_size: int) -> Tuple[Tensor, Tensor]:
        if not self.mixed_match:
            return src_ind, tgt_ind
        # hungarian_matcher.py:88-91 (the sort is made stable so that the device result is defined)
        tgt_ind = tgt_ind % gt_size if gt_size > 0 else tgt_ind
        tgt_ind, ind = tgt_ind.sort(stable=True)
        return src_ind[ind].view(-1), tgt_ind

    def _solve(self, mats: Sequence[Tensor]):
        self._raise_deferred()
        pairs, status = ops.lsap_solve(mats)
        self.last_status = status
        if self.check_status:
            bad = status.nonzero().flatten().tolist()   # synchronises: opt-in
            if bad:
                code = int(status[bad[0]])
                raise ValueError("cost matrix is infeasible" if code == 1 else "matrix contains invalid numeric entries")
        elif status.numel():
            # Failures must not stay silent (SciPy raises ValueError; a failed problem here returns -1 indices, which
            # the criterion would use as "last query / last target"), but a check per call would put a host sync back
            # into every decoder layer.  Deferred check: the worst status code travels to pinned host memory behind an
            # event, and the NEXT call (or ``check_deferred()``) raises once the copy has landed -- never blocking.
            if self._flag_host is None:
                self._flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
                self._flag_event = torch.cuda.Event()
            self._flag_host.copy_(status.max().reshape(1), non_blocking=True)
            self._flag_event.record()
            self._flag_pending = True
        return pairs

    def _raise_deferred(self, wait: bool = False) -> None:
        if not self._flag_pending:
            return
        if wait:
            self._flag_event.synchronize()
        elif not self._flag_event.query():
            return
        self._flag_pending = False
        code = int(self._flag_host[0])
        if code:
            raise ValueError(("cost matrix is infeasible" if code == 1 else "matrix contains invalid numeric entries")
                             + " (reported by an earlier HungarianMatcher call; its indices were -1)")

    def check_deferred(self) -> None:
        """Blocks until the status of the most recent call is known and raises ``ValueError`` if a problem failed
        (what SciPy raises at hungarian_matcher.py:80).  Call it once per step if failures must surface in-step."""
        self._raise_deferred(wait=True)

    def _costs(self, pred_boxes, pred_logits, gt_boxes, gt_labels) -> List[Tensor]:
        """Cost matrices of a batch of problems: one fused launch (``rdetr_match_cost``, the eager chain's
        values bit for bit) or, with ``fused_cost=False``, ``calculate_cost`` per problem as upstream."""
        if self.fused_cost:
            return ops.match_cost([b.float() for b in pred_boxes], [l.float() for l in pred_logits], [b.float() for b in gt_boxes],
                                  list(gt_labels), self.cost_class, self.cost_bbox, self.cost_giou, self.focal_alpha, self.focal_gamma)
        return [self.calculate_cost(pb, pl, gb, gl) for pb, pl, gb, gl in zip(pred_boxes, pred_logits, gt_boxes, gt_labels)]

    @torch.no_grad()
    def forward(self, pred_boxes: Tensor, pred_logits: Tensor, gt_boxes: Tensor, gt_labels: Tensor, gt_copy: int = 1):
        return self.match_batch([pred_boxes], [pred_logits], [gt_boxes], [gt_labels], gt_copy)[0]

    @torch.no_grad()
    def match_batch(self, pred_boxes: Sequence[Tensor], pred_logits: Sequence[Tensor], gt_boxes: Sequence[Tensor],
                    gt_labels: Sequence[Tensor], gt_copy: int = 1) -> List[Tuple[Tensor, Tensor]]:
        """Any number of (prediction set, image) problems in two launches (costs, assignment); same result as
        ``list(map(self, pred_boxes, pred_logits, gt_boxes, gt_labels))`` (set_criterion.py:126).  The problems
        need not share the number of queries, so the sets of several decoder layers can be matched together."""
        prepared = [self._prepare(c, gt_copy) for c in self._costs(pred_boxes, pred_logits, gt_boxes, gt_labels)]
        pairs = self._solve([c for c, _ in prepared])
        return [self._finish(s, t, g) for (s, t), (_, g) in zip(pairs, prepared)]


@torch.no_grad()
def match_prediction_sets(matcher: HungarianMatcher, outputs: dict, targets: Sequence[dict],
                          two_stage_binary_cls: bool = False) -> dict:
    """Match every prediction set of one criterion call -- the final layer, each ``aux_outputs[i]`` and
    ``enc_outputs`` -- against the images' targets with ONE cost launch and ONE solver launch.

    ``SetCriterion.forward`` of the reference (models/bricks/set_criterion.py:133-171) walks these sets one
    after the other and matches inside ``calculate_loss`` (:123-126), which already accepts precomputed
    ``indices``.  -> ``{"main": [...], "aux": [[...], ...], "enc": [...]}``, each list holding one
    ``(src_ind, tgt_ind)`` per image, ready to be passed as that argument.  ``two_stage_binary_cls`` zeroes the
    labels for the encoder set (:165-167)."""
    gt_boxes = [t["boxes"] for t in targets]
    gt_labels = [t["labels"] for t in targets]
    sets = [("main", outputs["pred_boxes"], outputs["pred_logits"], gt_labels)]
    for aux in outputs.get("aux_outputs", ()):
        sets.append(("aux", aux["pred_boxes"], aux["pred_logits"], gt_labels))
    if "enc_outputs" in outputs:
        enc = outputs["enc_outputs"]
        enc_labels = [torch.zeros_like(l) for l in gt_labels] if two_stage_binary_cls else gt_labels
        sets.append(("enc", enc["pred_boxes"], enc["pred_logits"], enc_labels))
    # the fused cost kernel takes one class count per launch: group the sets by it (binary encoder heads differ)
    result = {"main": None, "aux": [], "enc": None}
    by_classes = {}
    for entry in sets:
        by_classes.setdefault(entry[2].shape[-1], []).append(entry)
    solved = {}
    for group in by_classes.values():
        pb, pl, gb, gl = [], [], [], []
        for _, boxes, logits, labels in group:
            pb += list(boxes); pl += list(logits); gb += gt_boxes; gl += labels
        pairs = matcher.match_batch(pb, pl, gb, gl)
        n = len(targets)
        for k, entry in enumerate(group):
            solved[id(entry)] = pairs[k * n:(k + 1) * n]
    for entry in sets:
        if entry[0] == "aux":
            result["aux"].append(solved[id(entry)])
        else:
            result[entry[0]] = solved[id(entry)]
    return result
