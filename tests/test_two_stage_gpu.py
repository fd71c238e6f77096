"""Two-stage query selection on the B200 (csrc/topk.cu) against the oracle, the reference's
fixtures and torch's own expression (relation_transformer.py:90-96).  Index work: indices and gathered class rows are
bit-exact; the sigmoid of the selected boxes is torch's arithmetic (torch.equal on the same device)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import two_stage
from relation_detr_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference_expression(cls, box_unact, k):
    coord = box_unact.sigmoid()
    idx = torch.topk(cls.max(-1)[0], k, dim=1)[1].unsqueeze(-1)
    return cls.gather(1, idx.expand(-1, -1, cls.shape[-1])), coord.gather(1, idx.expand(-1, -1, 4)), idx.squeeze(-1)


@pytest.mark.parametrize("name", ["twostage_small", "twostage_pad"])
@pytest.mark.parametrize("head", ["main", "hybrid"])
def test_reference_fixtures(name, head):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    k = int(z["k_" + head])
    cls, box, idx = ops.two_stage_select(torch.from_numpy(z[head + "_class"]).to(DEV), torch.from_numpy(z[head + "_coord_unact"]).to(DEV), k)
    assert np.array_equal(cls.cpu().numpy(), z[f"expected_{head}_class"])
    assert np.abs(box.cpu().numpy() - z[f"expected_{head}_coord"]).max() <= 1e-6
    assert np.array_equal(idx.cpu().numpy(), two_stage.two_stage_select(z[head + "_class"], z[head + "_coord_unact"], k)[2])


@pytest.mark.parametrize("B,S,C,k", [(8, 22323, 91, 900), (8, 22323, 91, 1500), (2, 22323, 91, 900), (1, 204098, 91, 900),
                                     (3, 1000, 5, 1000), (2, 4097, 91, 4096), (5, 33, 1, 1), (1, 1, 3, 1),
                                     (3, 5000, 7, 300), (2, 10001, 3, 1500), (1, 400003, 2, 900)])   # clusters of 2 / 4 CTAs; keys not staged
def test_equals_torch_expression_bit_for_bit(B, S, C, k):
    g = torch.Generator(device=DEV).manual_seed(S + k)
    cls = torch.randn((B, S, C), device=DEV, generator=g) - 4.6     # the head's prior-probability bias: shared leading bytes
    box = torch.randn((B, S, 4), device=DEV, generator=g) * 3
    want_cls, want_box, want_idx = _reference_expression(cls, box, k)
    got_cls, got_box, got_idx = ops.two_stage_select(cls, box, k)
    scores = cls.max(-1)[0]
    if all(torch.unique(scores[b]).numel() == S for b in range(B)):   # tie-free: the very indices torch returns
        assert torch.equal(got_idx, want_idx)
        assert torch.equal(got_cls, want_cls) and torch.equal(got_box, want_box)
    else:                                                               # ties: same scores, rows consistent with our indices
        assert torch.equal(scores.gather(1, got_idx), scores.gather(1, want_idx))
        assert torch.equal(got_cls, cls.gather(1, got_idx[..., None].expand(-1, -1, C)))
    v, i = ops.topk_rows(scores, k)
    assert torch.equal(i, got_idx) and torch.equal(v, scores.gather(1, got_idx))


def test_ties_nan_and_infinities_follow_the_oracle():
    rng = np.random.default_rng(3)
    s = rng.integers(-3, 4, size=(4, 5000)).astype(np.float32)        # ~700 copies of every value
    s[0, 17] = np.nan
    s[0, 4000] = np.nan
    s[1, 5] = np.inf
    s[2, :] = -4.59512                                                 # a whole padded image: every score equal
    s[3, 100:200] = -np.inf
    s[3, 0] = -0.0
    for k in (1, 700, 701, 2048, 4096):
        idx = ops.topk_rows(torch.from_numpy(s).to(DEV), k)[1].cpu().numpy()
        assert np.array_equal(idx, two_stage.topk_rows(s, k)), k
    # the same through every cluster size (1, 2, 4, 8 CTAs per row) and the unstaged path: rows tiled to longer rows
    for reps, k in ((1, 900), (3, 900), (9, 4096), (70, 1500)):
        t = np.tile(s[:, :4000], (1, reps))
        idx = ops.topk_rows(torch.from_numpy(t).to(DEV), k)[1].cpu().numpy()
        assert np.array_equal(idx, two_stage.topk_rows(t, k)), (reps, k)


def test_errors_are_loud():
    s = torch.randn(2, 10, device=DEV)
    with pytest.raises(RuntimeError, match="out of range"):
        ops.topk_rows(s, 11)
    with pytest.raises(RuntimeError, match="supported maximum"):
        ops.topk_rows(torch.randn(1, 5000, device=DEV), 4097)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.topk_rows(torch.randn(2, 10), 3)


def test_gradients_equal_autograd_through_the_reference_expression():
    g = torch.Generator(device=DEV).manual_seed(5)
    cls = torch.randn((2, 3000, 91), device=DEV, generator=g, requires_grad=True)
    box = (torch.randn((2, 3000, 4), device=DEV, generator=g) * 2).requires_grad_(True)
    gc = torch.randn((2, 300, 91), device=DEV, generator=g)
    gb = torch.randn((2, 300, 4), device=DEV, generator=g)
    tc, tb, _ = ops.two_stage_select(cls, box, 300)
    ((tc * gc).sum() + (tb * gb).sum()).backward()
    got = (cls.grad.clone(), box.grad.clone())
    cls.grad = box.grad = None
    wc, wb, _ = _reference_expression(cls, box, 300)
    ((wc * gc).sum() + (wb * gb).sum()).backward()
    assert torch.equal(got[0], cls.grad)
    assert torch.equal(got[1], box.grad)
    # and against the oracle's adjoint in float64
    tc64, tb64, idx = two_stage.two_stage_select(cls.detach().cpu().double().numpy(), box.detach().cpu().double().numpy(), 300)
    d_cls, d_box = two_stage.two_stage_select_backward(gc.cpu().double().numpy(), gb.cpu().double().numpy(), tb64, idx, 3000)
    assert np.array_equal(got[0].cpu().numpy(), d_cls.astype(np.float32))
    assert np.abs(got[1].cpu().double().numpy() - d_box).max() <= 1e-6
