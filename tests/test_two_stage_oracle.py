"""The two-stage selection oracle (oracle/two_stage.py) against the reference's own outputs (tests/golden/twostage_*.npz,
written by oracle/make_golden_two_stage.py from RelationTransformer.forward) and against torch.topk.  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import two_stage


@pytest.mark.parametrize("name", ["twostage_small", "twostage_pad"])
@pytest.mark.parametrize("head", ["main", "hybrid"])
def test_oracle_reproduces_what_the_reference_forward_returned(name, head):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    k = int(z["k_" + head])
    cls, box, idx = two_stage.two_stage_select(z[head + "_class"], z[head + "_coord_unact"], k)
    assert np.array_equal(cls, z[f"expected_{head}_class"])           # index work: the gathered rows are the same bits
    assert np.abs(box - z[f"expected_{head}_coord"]).max() <= 1e-6    # numpy's exp vs ATen's, float32
    assert len({(b, i) for b in range(idx.shape[0]) for i in idx[b]}) == idx.size


def test_oracle_equals_torch_topk_on_tie_free_scores():
    rng = np.random.default_rng(0)
    for S, k in ((1, 1), (37, 37), (1000, 17), (22323, 900), (5000, 1500)):
        s = rng.permutation(S * 3).astype(np.float32).reshape(3, S) - S   # distinct values per row, both signs
        want = torch.topk(torch.from_numpy(s), k, dim=1)[1].numpy()
        assert np.array_equal(two_stage.topk_rows(s, k), want)


def test_oracle_tie_and_nan_conventions():
    s = np.array([[1.0, 3.0, 3.0, np.nan, -np.inf, 3.0, np.inf, 1.0]], np.float32)
    assert two_stage.topk_rows(s, 8)[0].tolist() == [3, 6, 1, 2, 5, 0, 7, 4]      # NaN, +inf, the 3s by index, the 1s by index, -inf
    vals = torch.topk(torch.from_numpy(s), 8, dim=1)[0].numpy()                     # torch agrees on the VALUES
    assert np.array_equal(np.nan_to_num(vals, nan=9e9), np.nan_to_num(s[0, two_stage.topk_rows(s, 8)[0]], nan=9e9)[None])
    with pytest.raises(ValueError):
        two_stage.topk_rows(s, 9)


def test_backward_is_the_adjoint_of_the_forward():
    rng = np.random.default_rng(1)
    cls, box = rng.standard_normal((2, 50, 5)), rng.standard_normal((2, 50, 4))
    tc, tb, idx = two_stage.two_stage_select(cls, box, 9)
    gc, gb = rng.standard_normal(tc.shape), rng.standard_normal(tb.shape)
    dcls, dbox = two_stage.two_stage_select_backward(gc, gb, tb, idx, 50)
    eps = 1e-6
    d = rng.standard_normal(box.shape)
    tb2 = two_stage.two_stage_select(cls, box + eps * d, 9)[1]
    assert abs(((tb2 - tb) * gb).sum() / eps - (dbox * d).sum()) <= 1e-5
    assert np.array_equal(np.take_along_axis(dcls, idx[..., None], 1), gc) and np.count_nonzero(dcls) == gc.size
