"""The reference's training step, restated without ``accelerate`` (absent from this image) so that the same loop
can run the unmodified reference model and the same model with the B200 operators installed.

Follows ``util/engine.py:46-60`` (loss = sum of the weighted loss dict the model returns, zero_grad, backward,
clip_grad_norm_(0.1), optimizer.step) with the optimizer of ``configs/train_config.py:41-46``: AdamW(lr 1e-4,
weight_decay 1e-4) over ``optimizer.param_dict.finetune_backbone_and_linear_projection``.  Mixed precision is
``torch.autocast(bfloat16)`` around the model call, which is what ``Accelerator(mixed_precision="bf16")`` does.
The per-step logging all-reduce + ``.item()`` of the reference (``engine.py:66-70``) is optional (``log_sync``).
"""
from __future__ import annotations

import contextlib


def build_optimizer(model, lr: float = 1e-4, weight_decay: float = 1e-4):
    import torch
    from optimizer import param_dict  # the reference's own grouping (lr x 0.1 for backbone / sampling_offsets, wd 0 for norm+bias)

    groups = param_dict.finetune_backbone_and_linear_projection(model, lr)
    return torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay, betas=(0.9, 0.999))


def train_step(model, images, targets, optimizer, autocast_dtype=None, max_norm: float = 0.1, log_sync: bool = False,
               autocast_cache: bool = True):
    """One optimisation step; returns the detached total loss (a 0-d tensor on the device).  ``autocast_cache=False`` is what
    CUDA-graph-captured parts need under autocast (relation_detr_b200.graphs)."""
    import torch

    ctx = torch.autocast("cuda", dtype=autocast_dtype, cache_enabled=autocast_cache) if autocast_dtype is not None else contextlib.nullcontext()
    with ctx:
        loss_dict = model(images, targets)
        losses = sum(loss for loss in loss_dict.values())
    optimizer.zero_grad()
    losses.backward()
    if max_norm > 0:
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
    optimizer.step()
    if log_sync:  # engine.py:66-70: reduce every loss over the ranks for logging, then .item()
        import torch.distributed as dist

        with torch.no_grad():
            red = {k: v.detach().clone() for k, v in loss_dict.items()}
            if dist.is_available() and dist.is_initialized():
                for v in red.values():
                    dist.all_reduce(v)
                    v /= dist.get_world_size()
            float(sum(red.values()).item())
    return losses.detach()
