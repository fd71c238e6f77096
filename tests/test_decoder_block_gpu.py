"""Integration: a relation decoder stack calling the two operators the way the reference's decoder
does (tools/decoder_harness.py), ours vs the same weights routed through the oracle."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def test_decoder_stack_matches_oracle_backed_stack():
    import decoder_harness as dh

    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # keep the oracle's 1x1 conv in fp32
    try:
        levels = ((25, 42), (13, 21), (7, 11), (4, 6))
        ours, oracle = dh.build_pair(0, layers=3)
        inp = dh.make_inputs(2, 60, 40, levels, seed=3)
        co, bo = ours(**inp)
        cr, br = oracle(**inp)
        assert torch.isfinite(co).all() and torch.isfinite(bo).all()
        assert (co - cr).abs().max().item() <= 2e-3 and (bo - br).abs().max().item() <= 2e-4
        (co.sum() + bo.sum()).backward()
        (cr.sum() + br.sum()).backward()
        worst = 0.0
        for (n, p), (_, q) in zip(ours.named_parameters(), oracle.named_parameters()):
            if q.grad is None:
                assert p.grad is None or p.grad.abs().max() == 0, n
                continue
            assert p.grad is not None, n
            den = max(q.grad.abs().max().item(), 1e-3)
            worst = max(worst, (p.grad - q.grad).abs().max().item() / den)
        assert worst <= 5e-3, worst
        # hybrid pass (no relation bias) and bf16 autocast run through the same code
        c2, b2 = ours(**dh.make_inputs(2, 90, 0, levels, seed=4), skip_relation=True)
        assert torch.isfinite(c2).all()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            c3, b3 = ours(**inp)
        assert torch.isfinite(c3.float()).all() and (b3.float() - bo).abs().max().item() <= 5e-2
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
