"""CUDA-graph capture of the static-shape parts of the reference's detector (partial-network capture).

The real training step is bound by the host, not by the GPU: 11 K kernel launches per step against 78 ms of device work in
bf16 (``profiles/r02g_train_host_profile_bf16.txt``).  The operators of this package never allocate behind torch's back,
never synchronise and read their shape tensors on the device, so every part of the model whose tensor shapes do not depend
on the data can be captured, forward AND backward, into CUDA graphs: the backbone, ``RelationTransformerEncoder`` and the
two ``RelationTransformerDecoder`` passes (main + hybrid) -- about 30 % of the step's launches.  What stays eager is
upstream's data-dependent glue: the denoising generator, the two-stage selection, the criterion (its ``.item()`` at
``set_criterion.py:147`` cannot be captured) and the optimizer.

    handle = graphs.capture_static_parts(model, images, targets, autocast_dtype=torch.bfloat16)
    ...train as before; ``handle.release()`` restores the eager forwards...

No reference file is edited: the capture uses ``torch.cuda.make_graphed_callables`` on thin positional wrappers around
the reference's own modules (which are called with keyword arguments upstream, ``relation_transformer.py:72-79, 125-146``)
and rebinds ``module.forward`` on the instances.  Constraints (those of partial-network capture): fixed input size and a
fixed number of denoising queries (the synthetic benchmark; production data needs one capture per shape bucket), the same
``requires_grad`` pattern on every call, and -- under autocast -- ``cache_enabled=False`` for the step (``autocast_kwargs``).
A call whose arguments do not match the captured signature runs the eager forward.

In eval mode (``model.eval()`` before the call) the same parts are captured forward-only (plain ``torch.cuda.CUDAGraph`` with
static input / output buffers): whole-model inference at batch 1 is launch-bound too.  The captured outputs live in static
buffers that the next call overwrites -- upstream consumes them inside the same forward (the post-processor builds new tensors).
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn

__all__ = ["capture_static_parts", "GraphHandle", "autocast_kwargs"]


def autocast_kwargs(dtype: Optional[torch.dtype]) -> dict:
    """Arguments of the ``torch.autocast`` context a step with captured parts must run under."""
    return {"device_type": "cuda", "dtype": dtype, "cache_enabled": False, "enabled": dtype is not None}


class _Positional(nn.Module):
    """``inner`` called with the keyword arguments ``names`` (tensors, positional here) plus the constants ``consts``."""

    def __init__(self, inner: nn.Module, names: Sequence[str], consts: Dict[str, object]):
        super().__init__()
        self.inner = inner
        self.names = tuple(names)
        self.consts = dict(consts)

    def forward(self, *args: Tensor):
        kw = dict(zip(self.names, args))
        kw.update(self.consts)
        return type(self.inner).forward(self.inner, **kw)   # the class's forward: the instance attribute is the dispatcher


class _Signature:
    def __init__(self, names, consts, tensors):
        self.names, self.consts = tuple(names), dict(consts)
        self.meta = tuple((t.shape, t.dtype, t.requires_grad) for t in tensors)

    def matches(self, kwargs) -> bool:
        tens = {k: v for k, v in kwargs.items() if isinstance(v, Tensor)}
        rest = {k: v for k, v in kwargs.items() if not isinstance(v, Tensor)}
        if tuple(sorted(tens)) != tuple(sorted(self.names)) or rest != self.consts:
            return False
        grad = torch.is_grad_enabled()
        return all((tens[n].shape, tens[n].dtype, tens[n].requires_grad and grad) == (m[0], m[1], m[2]) for n, m in zip(self.names, self.meta))


class _ForwardGraph:
    """Forward-only capture of ``fn(*tensors)`` (inference): static inputs, one graph, static outputs."""

    def __init__(self, fn, sample_args: Sequence[Tensor], num_warmup_iters: int = 3):
        self.static_in = [a.detach().clone() for a in sample_args]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(num_warmup_iters):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *args: Tensor):
        for dst, src in zip(self.static_in, args):
            dst.copy_(src)
        self.graph.replay()
        return self.static_out


class GraphHandle:
    """What ``capture_static_parts`` returns: the captured parts and the way back."""

    def __init__(self):
        self.parts: List[str] = []
        self._undo = []

    def release(self) -> None:
        for fn in reversed(self._undo):
            fn()
        self._undo.clear()
        self.parts.clear()


def _sample(t: Tensor) -> Tensor:
    s = t.detach().clone()
    if t.requires_grad:
        s.requires_grad_(True)
    return s


def _record_calls(model, modules: Dict[str, nn.Module], images, targets, ctx) -> Dict[str, List[Tuple[tuple, dict]]]:
    calls: Dict[str, List[Tuple[tuple, dict]]] = {k: [] for k in modules}
    hooks = []
    for name, mod in modules.items():
        def pre(m, args, kwargs, name=name):
            # detached samples only: a reference to the live tensors would keep the recording pass's autograd graph -- and with it
            # AccumulateGrad nodes bound to the default stream -- alive, which invalidates the capture of the backward
            calls[name].append((tuple(_sample(a) if isinstance(a, Tensor) else a for a in args),
                                {k: (_sample(v) if isinstance(v, Tensor) else v) for k, v in kwargs.items()}))
        hooks.append(mod.register_forward_pre_hook(pre, with_kwargs=True))
    try:
        with ctx():
            out = model(images, targets)
        del out
    finally:
        for h in hooks:
            h.remove()
    import gc

    gc.collect()
    return calls


def capture_static_parts(model: nn.Module, images, targets, autocast_dtype: Optional[torch.dtype] = None,
                         parts: Sequence[str] = ("backbone", "encoder", "decoder"), num_warmup_iters: int = 3) -> GraphHandle:
    """Captures the listed parts of a reference ``RelationDETR`` in training mode for the input shapes of ``(images, targets)``.
    Call it BEFORE wrapping the model in DistributedDataParallel."""
    training = model.training   # training: forward + backward graphs (make_graphed_callables); eval: forward-only graphs
    handle = GraphHandle()

    def ctx():
        return torch.autocast(**autocast_kwargs(autocast_dtype)) if autocast_dtype is not None else contextlib.nullcontext()

    transformer = model.transformer
    modules = {}
    if "backbone" in parts:
        modules["backbone"] = model.backbone
    if "encoder" in parts:
        modules["encoder"] = transformer.encoder
    if "decoder" in parts:
        modules["decoder"] = transformer.decoder
    if training:
        calls = _record_calls(model, modules, images, targets, ctx)
    else:
        with torch.no_grad():
            calls = _record_calls(model, modules, images, None, ctx)
    torch.cuda.synchronize()

    with ctx():
        if "backbone" in modules and calls["backbone"]:
            (x,), kw = calls["backbone"][0]
            if not kw and isinstance(x, Tensor):
                bb = model.backbone
                had_bb = "forward" in bb.__dict__
                eager = bb.forward
                if training:
                    torch.cuda.make_graphed_callables(bb, (x,), num_warmup_iters=num_warmup_iters)   # patches bb.forward in place
                else:
                    fg = _ForwardGraph(lambda t, bb=bb: type(bb).forward(bb, t), (x,), num_warmup_iters)
                    sig = (x.shape, x.dtype)

                    def bb_dispatch(t, bb=bb, fg=fg, sig=sig):
                        if not bb.training and (t.shape, t.dtype) == sig and not torch.is_grad_enabled():
                            return fg(t)
                        return type(bb).forward(bb, t)
                    bb.forward = bb_dispatch
                handle.parts.append("backbone")

                def undo_bb(bb=bb, had_bb=had_bb, eager=eager):
                    if had_bb:
                        bb.forward = eager
                    else:
                        bb.__dict__.pop("forward", None)
                handle._undo.append(undo_bb)
        for name in ("encoder", "decoder"):
            if name not in modules:
                continue
            inner = modules[name]
            variants = []
            for args, kw in calls[name]:
                if args:   # upstream calls both with keywords only; anything else stays eager
                    continue
                names = [k for k, v in kw.items() if isinstance(v, Tensor)]
                consts = {k: v for k, v in kw.items() if not isinstance(v, Tensor)}
                tensors = [kw[k] for k in names]
                wrapper = _Positional(inner, names, consts)
                wrapper.train(training)
                if training:
                    torch.cuda.make_graphed_callables(wrapper, tuple(tensors), num_warmup_iters=num_warmup_iters,
                                                      allow_unused_input=True)
                    variants.append((_Signature(names, consts, tensors), wrapper))
                else:
                    variants.append((_Signature(names, consts, tensors), _ForwardGraph(wrapper, tensors, num_warmup_iters)))
            if not variants:
                continue

            def dispatch(*args, _inner=inner, _variants=variants, _training=training, **kwargs):
                if not args and _inner.training == _training and (_training or not torch.is_grad_enabled()):
                    for sig, wrapper in _variants:
                        if sig.matches(kwargs):
                            return wrapper(*[kwargs[n] for n in sig.names])
                return type(_inner).forward(_inner, *args, **kwargs)

            had = "forward" in inner.__dict__
            old = inner.__dict__.get("forward")
            inner.forward = dispatch
            handle.parts.append(f"{name} x{len(variants)}")

            def undo(inner=inner, had=had, old=old):
                if had:
                    inner.forward = old
                else:
                    del inner.__dict__["forward"]
            handle._undo.append(undo)
    torch.cuda.synchronize()
    return handle
