"""CUDA-graph capture of the static parts of the reference's detector (relation_detr_b200/graphs.py) on the B200: the captured
backbone / encoder / decoder passes must give the loss dict and the parameter gradients of the eager model."""
import copy
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import refmodel  # noqa: E402
from relation_detr_b200 import graphs  # noqa: E402
from relation_detr_b200 import install as rinstall  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refmodel.available(), reason="baseline/_ref not installed")]
DEV = "cuda:0"


def _step(model, images, targets, amp):
    model.zero_grad(set_to_none=True)
    torch.manual_seed(123)   # the denoising generator draws its noise inside the forward
    with torch.autocast("cuda", dtype=amp, cache_enabled=False, enabled=amp is not None):
        loss_dict = model(copy.deepcopy(images), copy.deepcopy(targets))
        loss = sum(loss_dict.values())
    loss.backward()
    return ({k: v.detach().double().item() for k, v in loss_dict.items()},
            {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None})


@pytest.mark.parametrize("amp", [None, torch.bfloat16])
def test_captured_parts_reproduce_the_eager_step(amp):
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        refmodel.activate()
        rinstall.uninstall()
        rinstall.install()
        torch.manual_seed(0)
        model, _ = refmodel.build_relation_detr_r50(enc_layers=2, dec_layers=2)
        rinstall.uninstall()
        model = model.to(DEV).train()
        images, targets = refmodel.synthetic_batch(2, DEV, seed=1, height=416, width=544, boxes_per_image=5)
        eager_losses, eager_grads = _step(model, images, targets, amp)
        handle = graphs.capture_static_parts(model, images, targets, autocast_dtype=amp)
        try:
            assert handle.parts == ["backbone", "encoder x1", "decoder x2"], handle.parts
            for _ in range(2):   # replay twice: static buffers are reused
                losses, grads = _step(model, images, targets, amp)
            # different images of the same size: the graphs must read their inputs, not the sample
            images2, targets2 = refmodel.synthetic_batch(2, DEV, seed=7, height=416, width=544, boxes_per_image=5)
            losses2, _ = _step(model, images2, targets2, amp)
        finally:
            handle.release()
        ref_losses2, _ = _step(model, images2, targets2, amp)
        tol = 2e-2 if amp is not None else 2e-4
        assert losses.keys() == eager_losses.keys()
        for k in eager_losses:
            assert abs(losses[k] - eager_losses[k]) <= tol * max(1.0, abs(eager_losses[k])), (k, losses[k], eager_losses[k])
            assert abs(losses2[k] - ref_losses2[k]) <= tol * max(1.0, abs(ref_losses2[k])), (k, losses2[k], ref_losses2[k])
        assert grads.keys() == eager_grads.keys()
        num = sum(((grads[n].double() - eager_grads[n].double()) ** 2).sum() for n in grads).sqrt().item()
        den = sum((eager_grads[n].double() ** 2).sum() for n in grads).sqrt().item()
        assert num / den <= (5e-2 if amp is not None else 5e-3), num / den
        assert "forward" not in model.transformer.encoder.__dict__ and "forward" not in model.transformer.decoder.__dict__
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        rinstall.uninstall()


@pytest.mark.parametrize("amp", [None, torch.bfloat16])
def test_forward_only_capture_reproduces_eager_inference(amp):
    """model.eval(): backbone, encoder and the decoder pass captured forward-only; detections must be those of the eager model,
    for the captured image and for another image of the same size."""
    refmodel.activate()
    rinstall.uninstall()
    rinstall.install()
    torch.manual_seed(0)
    model, _ = refmodel.build_relation_detr_r50(enc_layers=2, dec_layers=2)
    rinstall.uninstall()
    model.eval_transform = None
    model = model.to(DEV).eval()
    g = torch.Generator(device=DEV).manual_seed(5)
    img_a = torch.randn((3, 416, 544), device=DEV, generator=g)
    img_b = torch.randn((3, 416, 544), device=DEV, generator=g)

    def detect(img):
        with torch.inference_mode(), torch.autocast("cuda", dtype=amp, enabled=amp is not None, cache_enabled=False):
            det = model((img,))[0]
        return {k: v.clone() for k, v in det.items()}

    want_a, want_b = detect(img_a), detect(img_b)
    handle = graphs.capture_static_parts(model, (img_a,), None, autocast_dtype=amp)
    try:
        assert handle.parts == ["backbone", "encoder x1", "decoder x1"], handle.parts
        got_a, got_b, got_a2 = detect(img_a), detect(img_b), detect(img_a)
    finally:
        handle.release()
    for got, want in ((got_a, want_a), (got_b, want_b), (got_a2, want_a)):
        assert got.keys() == want.keys()
        for k in want:
            if want[k].is_floating_point():
                assert (got[k].float() - want[k].float()).abs().max().item() <= 1e-4 * max(1.0, want[k].float().abs().max().item()), k
            else:
                assert torch.equal(got[k], want[k]), k
    assert not (want_a["scores"] == want_b["scores"]).all()   # the two images do give different detections
    assert "forward" not in model.backbone.__dict__ and "forward" not in model.transformer.decoder.__dict__
