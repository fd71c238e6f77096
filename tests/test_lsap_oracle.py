"""CPU: the LSAP oracle (oracle/lsap_oracle.c) against SciPy -- the third-party solver the reference's
matcher calls (hungarian_matcher.py:80,87) -- and against index fixtures produced by the reference's
HungarianMatcher itself (tests/golden/matcher_*.npz, oracle/make_golden_matcher.py)."""
import glob
import os

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from oracle import c_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _matrix(kind, nr, nc, rng):
    if kind == "float":
        return rng.random((nr, nc)).astype(np.float32)
    if kind == "small_int":                       # many equal reduced costs: exercises the tie rules
        return rng.integers(0, 4, (nr, nc)).astype(np.float64)
    if kind == "dup_cols":                        # hybrid branch: every target repeated (relation_detr.py:131)
        g = max(1, nc // 3)
        return np.tile(rng.random((nr, g)).astype(np.float32), (1, 3))
    if kind == "dup_rows":
        g = max(1, nr // 2)
        return np.tile(rng.random((g, nc)).astype(np.float32), (2, 1))
    if kind == "constant":
        return np.zeros((nr, nc))
    if kind == "negative":
        return (rng.random((nr, nc)) - 0.5).astype(np.float32) * 7
    if kind == "with_inf":
        c = rng.integers(0, 3, (nr, nc)).astype(np.float64)
        c[rng.random((nr, nc)) < 0.2] = np.inf
        return c
    raise AssertionError(kind)


def _both(c):
    def run(fn):
        try:
            return fn(c), None
        except ValueError as e:
            return None, str(e)
    return run(linear_sum_assignment), run(c_oracle.lsap)


@pytest.mark.parametrize("kind", ["float", "small_int", "dup_cols", "dup_rows", "constant", "negative", "with_inf"])
def test_oracle_returns_scipys_pairs(kind):
    rng = np.random.default_rng(sum(map(ord, kind)))
    for _ in range(300):
        nr, nc = int(rng.integers(1, 48)), int(rng.integers(1, 48))
        (ref, ref_err), (got, got_err) = _both(_matrix(kind, nr, nc, rng))
        assert ref_err == got_err
        if ref is not None:
            assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1]), (kind, nr, nc)


def test_oracle_at_matcher_sizes():
    rng = np.random.default_rng(7)
    for nq, ng, rep in ((900, 40, 1), (1500, 30, 6), (300, 0, 1), (25, 60, 1)):
        c = np.tile(rng.random((nq, ng)).astype(np.float32), (1, rep))
        ref, got = linear_sum_assignment(c), c_oracle.lsap(c)
        assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])


def test_oracle_rejects_what_scipy_rejects():
    for bad in (np.nan, -np.inf):
        c = np.ones((4, 6))
        c[2, 3] = bad
        with pytest.raises(ValueError, match="invalid numeric"):
            linear_sum_assignment(c)
        with pytest.raises(ValueError, match="invalid numeric"):
            c_oracle.lsap(c)
    c = np.full((3, 3), np.inf)
    with pytest.raises(ValueError, match="infeasible"):
        linear_sum_assignment(c)
    with pytest.raises(ValueError, match="infeasible"):
        c_oracle.lsap(c)


def finish_like_reference(src, tgt, gt_size, mixed):
    """hungarian_matcher.py:88-91 on numpy arrays (stable sort: the device path's definition)."""
    if not mixed:
        return src, tgt
    tgt = tgt % gt_size
    order = np.argsort(tgt, kind="stable")
    return src[order], tgt[order]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "matcher_*.npz"))), ids=os.path.basename)
def test_oracle_reproduces_the_reference_matcher(path):
    z = np.load(path)
    cost, mixed, gt_copy = z["cost"], bool(z["mixed"]), int(z["gt_copy"])
    gt_size = cost.shape[1]
    if mixed:
        gt_copy = min(int(cost.shape[0] * 0.5 / gt_size), gt_copy) if gt_size > 0 else gt_copy
        cost = np.tile(cost, (1, gt_copy))
    src, tgt = c_oracle.lsap(cost)
    src, tgt = finish_like_reference(src, tgt, gt_size, mixed)
    if mixed:   # upstream's sort is not declared stable: compare as a set of pairs and as sorted targets
        assert np.array_equal(tgt, z["tgt_ind"])
        assert sorted(zip(src.tolist(), tgt.tolist())) == sorted(zip(z["src_ind"].tolist(), z["tgt_ind"].tolist()))
    else:
        assert np.array_equal(src, z["src_ind"]) and np.array_equal(tgt, z["tgt_ind"])


def test_oracle_is_optimal_against_brute_force():
    """Independent of SciPy: on small matrices the oracle's total cost equals the minimum over all assignments."""
    import itertools
    rng = np.random.default_rng(3)
    for _ in range(200):
        nr, nc = int(rng.integers(1, 6)), int(rng.integers(1, 7))
        c = rng.integers(-5, 6, (nr, nc)).astype(np.float64) if rng.random() < 0.5 else rng.standard_normal((nr, nc))
        r, k = c_oracle.lsap(c)
        assert len(r) == min(nr, nc) and len(set(r.tolist())) == len(r) and len(set(k.tolist())) == len(k)
        if nr <= nc:
            best = min(sum(c[i, p[i]] for i in range(nr)) for p in itertools.permutations(range(nc), nr))
        else:
            best = min(sum(c[p[j], j] for j in range(nc)) for p in itertools.permutations(range(nr), nc))
        assert abs(c[r, k].sum() - best) <= 1e-12 * max(1.0, abs(best))


def test_eager_cost_terms_reproduce_the_reference_fixtures():
    """relation-detr_b200/matcher.py restates torchvision's box conversion and GIoU so that the package does not
    import torchvision; on the CPU its calculate_cost must give the bits the reference's matcher stored."""
    import torch
    import relation_detr_b200 as rd
    m = rd.HungarianMatcher(cost_class=2, cost_bbox=5, cost_giou=2, fused_cost=False)
    for path in sorted(glob.glob(os.path.join(GOLDEN, "matcher_*.npz"))):
        z = np.load(path)
        c = m.calculate_cost(*[torch.as_tensor(z[k]) for k in ("pred_boxes", "pred_logits", "gt_boxes", "gt_labels")])
        assert np.array_equal(c.numpy(), z["cost"]), os.path.basename(path)


def test_oracle_fuzz_with_hypothesis():
    """Property: for any finite matrix (small integers, so that equal reduced costs are the rule, not the exception)
    the oracle returns SciPy's pairs."""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    from hypothesis.extra.numpy import arrays

    shapes = st.tuples(st.integers(1, 9), st.integers(1, 9))

    @settings(max_examples=300, deadline=None)
    @given(shapes.flatmap(lambda s: arrays(np.float64, s, elements=st.integers(-2, 2).map(float))))
    def check(c):
        rr, cc = linear_sum_assignment(c)
        r, k = c_oracle.lsap(c)
        assert np.array_equal(r, rr) and np.array_equal(k, cc)

    check()
