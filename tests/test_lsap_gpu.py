"""GPU: rdetr_lsap_solve / HungarianMatcher against SciPy (the solver the reference calls,
hungarian_matcher.py:80,87), the C oracle and the fixtures produced by the reference's matcher.
Index work: the bar is bit-exact."""
import glob
import os

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

import relation_detr_b200 as rd
from oracle import c_oracle
from relation_detr_b200 import ops
from test_lsap_oracle import _matrix, finish_like_reference

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def _solve(mats):
    pairs, status = ops.lsap_solve([torch.as_tensor(np.asarray(m, dtype=np.float32)).to(DEV) for m in mats])
    torch.cuda.synchronize()
    return [(r.cpu().numpy(), c.cpu().numpy()) for r, c in pairs], status.cpu().numpy()


@pytest.mark.parametrize("kind", ["float", "small_int", "dup_cols", "dup_rows", "constant", "negative"])
def test_solver_returns_scipys_pairs(kind):
    rng = np.random.default_rng(sum(map(ord, kind)) + 1)
    mats = [_matrix(kind, int(rng.integers(1, 70)), int(rng.integers(1, 70)), rng).astype(np.float32) for _ in range(150)]
    got, status = _solve(mats)          # 150 problems: three launches of <= 64 CTAs
    assert not status.any()
    for m, (r, c) in zip(mats, got):
        rr, cc = linear_sum_assignment(m)
        assert np.array_equal(r, rr) and np.array_equal(c, cc), (kind, m.shape)


def test_solver_matches_the_oracle_with_infinite_entries_and_reports_status():
    rng = np.random.default_rng(11)
    mats = [_matrix("with_inf", int(rng.integers(1, 30)), int(rng.integers(1, 30)), rng).astype(np.float32) for _ in range(120)]
    nan = np.ones((5, 7), dtype=np.float32); nan[1, 2] = np.nan
    ninf = np.ones((7, 5), dtype=np.float32); ninf[6, 4] = -np.inf
    mats += [nan, ninf, np.full((3, 3), np.inf, dtype=np.float32)]
    got, status = _solve(mats)
    seen = set()
    for m, (r, c), st in zip(mats, got, status):
        try:
            rr, cc = c_oracle.lsap(m)
            assert st == 0 and np.array_equal(r, rr) and np.array_equal(c, cc)
        except ValueError as e:
            want = 1 if "infeasible" in str(e) else 2
            assert st == want and (r == -1).all() and (c == -1).all()
            seen.add(want)
    assert seen == {1, 2}


def test_solver_at_the_matchers_full_sizes():
    """900 queries x crowded image, and the hybrid branch's 1500 queries x 6 copies of every target
    (relation_detr.py:131-132): duplicated columns make the optimum non-unique, SciPy's choice must be kept."""
    rng = np.random.default_rng(5)
    mats = [rng.random((900, 100)).astype(np.float32), np.tile(rng.random((1500, 90)).astype(np.float32), (1, 6)),
            rng.random((300, 300)).astype(np.float32), rng.random((60, 1500)).astype(np.float32)]
    got, status = _solve(mats)
    assert not status.any()
    for m, (r, c) in zip(mats, got):
        rr, cc = linear_sum_assignment(m)
        assert np.array_equal(r, rr) and np.array_equal(c, cc), m.shape
        assert len(np.unique(c)) == len(c)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "matcher_*.npz"))), ids=os.path.basename)
def test_matcher_reproduces_the_reference_fixtures(path):
    z = np.load(path)
    mixed, gt_copy = bool(z["mixed"]), int(z["gt_copy"])
    m = rd.HungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, mixed_match=mixed)
    args = [torch.as_tensor(z[k]).to(DEV) for k in ("pred_boxes", "pred_logits", "gt_boxes", "gt_labels")]
    # (1) the solver on the reference's own cost matrix: exact
    cost = z["cost"]
    gt_size = cost.shape[1]
    if mixed:
        k = min(int(cost.shape[0] * 0.5 / gt_size), gt_copy) if gt_size > 0 else gt_copy
        cost = np.tile(cost, (1, k))
    (pair,), status = _solve([cost])
    src, tgt = finish_like_reference(pair[0], pair[1], gt_size, mixed)
    assert status[0] == 0
    if mixed:
        assert np.array_equal(tgt, z["tgt_ind"])
        assert sorted(zip(src.tolist(), tgt.tolist())) == sorted(zip(z["src_ind"].tolist(), z["tgt_ind"].tolist()))
    else:
        assert np.array_equal(src, z["src_ind"]) and np.array_equal(tgt, z["tgt_ind"])
    # (2) the whole matcher, cost computed on the device as upstream does
    s, t = m(*args, gt_copy=gt_copy)
    assert s.is_cuda and s.dtype == torch.int64 and t.dtype == torch.int64
    assert sorted(zip(s.tolist(), t.tolist())) == sorted(zip(z["src_ind"].tolist(), z["tgt_ind"].tolist()))
    c_dev = m.calculate_cost(*args).cpu().numpy()
    assert np.allclose(c_dev, z["cost"], rtol=1e-5, atol=1e-5)


def _ulps(a, b):
    ia, ib = a.contiguous().view(torch.int32).long(), b.contiguous().view(torch.int32).long()
    return int((ia - ib).abs().max()) if a.numel() else 0


def test_fused_cost_kernel_returns_the_eager_chains_bits():
    """rdetr_match_cost against HungarianMatcher.calculate_cost evaluated by eager torch on the same device
    (the reference's way, hungarian_matcher.py:40-72).  The assignment that follows must be the reference's,
    so the bar is equality of the float32 bit patterns, not a tolerance."""
    g = torch.Generator().manual_seed(9)
    eager = rd.HungarianMatcher(2, 5, 2, fused_cost=False)
    sets = []
    for nq, ng, ncls in ((900, 37, 91), (1500, 6 * 15, 91), (300, 1, 91), (50, 0, 91), (20, 64, 20), (7, 3, 2)):
        pb = torch.cat([torch.rand(nq, 2, generator=g) * 0.8 + 0.1, torch.rand(nq, 2, generator=g) * 0.3 + 0.02], -1).to(DEV)
        pl = (torch.randn(nq, ncls, generator=g) * 3 - 2).to(DEV)
        gb = torch.cat([torch.rand(ng, 2, generator=g) * 0.8 + 0.1, torch.rand(ng, 2, generator=g) * 0.3 + 0.02], -1).to(DEV)
        gl = torch.randint(0, ncls, (ng,), generator=g).to(DEV)
        sets.append((pb, pl, gb, gl))
    for ncls in (91, 20, 2):
        group = [s for s in sets if s[1].shape[1] == ncls]
        fused = ops.match_cost(*zip(*group), 2, 5, 2, 0.25, 2.0)
        for (pb, pl, gb, gl), c in zip(group, fused):
            want = eager.calculate_cost(pb, pl, gb, gl)
            assert c.shape == want.shape
            assert torch.equal(c, want), f"{tuple(c.shape)}: differs by up to {_ulps(c, want)} ulp, max abs {float((c - want).abs().max()):.3e}"
    # extreme logits (saturated sigmoid) and degenerate boxes keep the same bits too
    pb = torch.tensor([[0.5, 0.5, 0.2, 0.2], [0.1, 0.9, 1e-4, 0.3], [0.5, 0.5, 1.0, 1.0]], device=DEV)
    pl = torch.tensor([[40.0, -40.0, 0.0], [-100.0, 100.0, 1e-3], [15.9, -15.9, 7.0]], device=DEV)
    gb = torch.tensor([[0.5, 0.5, 0.2, 0.2], [0.9, 0.1, 0.05, 0.05]], device=DEV)
    gl = torch.tensor([0, 1], device=DEV)
    (c,) = ops.match_cost([pb], [pl], [gb], [gl], 2, 5, 2, 0.25, 2.0)
    want = eager.calculate_cost(pb, pl, gb, gl)
    assert torch.equal(c, want), f"differs by up to {_ulps(c, want)} ulp"


def test_match_batch_equals_per_image_calls_and_does_not_synchronise():
    g = torch.Generator().manual_seed(3)
    B, nq = 6, 900
    pb = (torch.rand(B, nq, 4, generator=g) * 0.5 + 0.1).to(DEV)
    pl = torch.randn(B, nq, 91, generator=g).to(DEV)
    gts = [int(n) for n in (0, 1, 7, 23, 64, 5)]
    gb = [(torch.rand(n, 4, generator=g) * 0.5 + 0.1).to(DEV) for n in gts]
    gl = [torch.randint(0, 91, (n,), generator=g).to(DEV) for n in gts]
    m = rd.HungarianMatcher(2, 5, 2)
    one = list(map(m, pb, pl, gb, gl))                       # set_criterion.py:126
    torch.cuda.synchronize()
    import warnings
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        torch.cuda.set_sync_debug_mode("warn")               # any implicit device->host sync is reported
        try:
            many = m.match_batch(pb, pl, gb, gl)
        finally:
            torch.cuda.set_sync_debug_mode("default")
    assert not [str(w.message) for w in caught if "called a synchronizing" in str(w.message)]
    for (s1, t1), (s2, t2), n in zip(one, many, gts):
        assert len(s1) == n and torch.equal(s1, s2) and torch.equal(t1, t2)
    for b, n in enumerate(gts):                              # against SciPy on the same device-computed cost
        rr, cc = linear_sum_assignment(m.calculate_cost(pb[b], pl[b], gb[b], gl[b]).cpu().numpy())
        assert np.array_equal(many[b][0].cpu().numpy(), rr) and np.array_equal(many[b][1].cpu().numpy(), cc)


def test_all_prediction_sets_of_a_criterion_call_in_one_go():
    """match_prediction_sets == the reference's walk over main / aux / enc outputs (set_criterion.py:133-171)."""
    from relation_detr_b200.matcher import match_prediction_sets
    g = torch.Generator().manual_seed(21)
    B, nq, ncls = 2, 300, 91
    mk = lambda: {"pred_boxes": (torch.rand(B, nq, 4, generator=g) * 0.5 + 0.1).to(DEV),  # noqa: E731
                  "pred_logits": torch.randn(B, nq, ncls, generator=g).to(DEV)}
    outputs = mk()
    outputs["aux_outputs"] = [mk() for _ in range(5)]
    outputs["enc_outputs"] = mk()
    targets = [{"boxes": (torch.rand(n, 4, generator=g) * 0.5 + 0.1).to(DEV), "labels": torch.randint(0, ncls, (n,), generator=g).to(DEV)}
               for n in (9, 0)]
    m = rd.HungarianMatcher(2, 5, 2)
    got = match_prediction_sets(m, outputs, targets, two_stage_binary_cls=True)

    def walk(o, labels):
        out = []
        for b in range(B):
            c = m.calculate_cost(o["pred_boxes"][b], o["pred_logits"][b], targets[b]["boxes"], labels[b]).cpu().numpy()
            out.append(linear_sum_assignment(c))
        return out

    labels = [t["labels"] for t in targets]
    for want, have in ((walk(outputs, labels), got["main"]), (walk(outputs["enc_outputs"], [torch.zeros_like(l) for l in labels]), got["enc"]),
                       *((walk(a, labels), h) for a, h in zip(outputs["aux_outputs"], got["aux"]))):
        for (r, c), (s, t) in zip(want, have):
            assert np.array_equal(r, s.cpu().numpy()) and np.array_equal(c, t.cpu().numpy())
    assert len(got["aux"]) == 5


def test_errors_are_raised_not_swallowed():
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.lsap_solve([torch.zeros(3, 3)])
    with pytest.raises(RuntimeError, match="float32"):
        ops.lsap_solve([torch.zeros(3, 3, dtype=torch.float64, device=DEV)])
    with pytest.raises(RuntimeError, match="shared memory"):
        ops.lsap_solve([torch.zeros(8000, 8000, device=DEV)])
    m = rd.HungarianMatcher(1, 1, 1, check_status=True)
    bad = torch.full((4, 4), float("nan"), device=DEV)
    with pytest.raises(ValueError, match="invalid numeric"):
        m._solve([bad])


def test_out_of_range_labels_poison_the_cost_and_failures_surface_without_a_sync_per_call():
    """a target label outside [0, num_classes) must not become an out-of-bounds read, and a failed problem
    must not stay silent.  The fused cost kernel writes NaN for such a column, the solver reports status 2, and the
    matcher raises ValueError -- immediately with check_status=True, otherwise on the next call / check_deferred()."""
    g = torch.Generator().manual_seed(0)
    pb = torch.cat([torch.rand(50, 2, generator=g) * 0.8 + 0.1, torch.rand(50, 2, generator=g) * 0.3 + 0.02], -1).to(DEV)
    pl = torch.randn(50, 80, generator=g).to(DEV)
    gb = torch.cat([torch.rand(4, 2, generator=g) * 0.8 + 0.1, torch.rand(4, 2, generator=g) * 0.3 + 0.02], -1).to(DEV)
    bad_labels = torch.tensor([3, 90, 7, 2], device=DEV)  # 90 >= 80 classes (COCO ids against an 80-class head)
    cost = ops.match_cost([pb], [pl], [gb], [bad_labels], 2.0, 5.0, 2.0, 0.25, 2.0)[0]
    assert torch.isnan(cost[:, 1]).all() and torch.isfinite(cost[:, [0, 2, 3]]).all()
    with pytest.raises(ValueError, match="invalid numeric"):
        rd.HungarianMatcher(2, 5, 2, check_status=True)(pb, pl, gb, bad_labels)
    m = rd.HungarianMatcher(2, 5, 2)
    src, tgt = m(pb, pl, gb, bad_labels)  # does not raise and does not synchronise ...
    assert int(m.last_status[0]) == 2 and bool((src < 0).all())
    with pytest.raises(ValueError, match="earlier HungarianMatcher call"):
        m.check_deferred()  # ... but the failure is reported
    good = torch.tensor([3, 9, 7, 2], device=DEV)
    m(pb, pl, gb, bad_labels)
    torch.cuda.synchronize()
    with pytest.raises(ValueError):
        m(pb, pl, gb, good)  # the next call reports the earlier failure once its status has landed
    src, tgt = m(pb, pl, gb, good)
    m.check_deferred()
    assert bool((src >= 0).all())
