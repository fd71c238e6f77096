"""The two hot-path operators as ``torch.library`` custom ops over the C ABI.

Mirrors the reference's operator interface for this path:

* ``MultiScaleDeformableAttnFunction.apply(value, value_spatial_shapes, level_start_index,
  sampling_locations, attention_weights, im2col_step)`` -- reference
  ``models/bricks/ms_deform_attn.py:35-84`` (pybind ``_C.ms_deform_attn_forward/backward``,
  ``ops/cuda/ms_deform_attn_cuda.cu:148-151``).
* ``position_relation_bias(src_boxes, tgt_boxes, weight, bias, ...)`` -- the fused body of
  ``PositionRelationEmbedding.forward`` (``models/bricks/relation_transformer.py:520-532``).

PyTorch is plumbing here (device memory, streams, autograd registration); the arithmetic is in
``csrc/*.cu``.  Nothing in this module computes on the CPU: non-CUDA tensors raise.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

__all__ = [
    "msda_forward", "msda_backward", "ms_deform_attn", "MultiScaleDeformableAttnFunction",
    "msda_fused_forward", "msda_fused_backward", "ms_deform_attn_fused",
    "relation_attention_forward", "relation_attention_backward", "relation_attention",
    "memory_fusion_forward", "memory_fusion_linear",
    "topk_rows", "two_stage_select_forward", "two_stage_select_backward", "two_stage_select",
    "relation_forward", "relation_backward", "position_relation_bias", "relation_dim_t",
    "lsap_solve", "match_cost",
]


def _ptr(t: Optional[Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _stream(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _value_dtype_code(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.DTYPE_F32
    if t.dtype == torch.bfloat16:
        return _lib.DTYPE_BF16
    raise RuntimeError(f"rdetr::msda: value dtype {t.dtype} unsupported (float32 or bfloat16)")


def _check_msda_inputs(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    # same preconditions the reference asserts (ms_deform_attn_cuda.cu:20-30), raised as RuntimeError
    for name, t in (("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                    ("sampling_loc", sampling_locations), ("attn_weight", attention_weights)):
        _require(t.is_cuda, f"{name} must be a CUDA tensor")
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
    _require(value.dim() == 4, "value must be [B, S, M, D]")
    _require(sampling_locations.dim() == 6 and sampling_locations.shape[-1] == 2, "sampling_loc must be [B, Nq, M, L, P, 2]")
    _require(attention_weights.shape == sampling_locations.shape[:-1], "attn_weight must be [B, Nq, M, L, P]")
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64,
             "spatial_shapes / level_start_index must be int64")
    _require(sampling_locations.dtype == torch.float32 and attention_weights.dtype == torch.float32,
             "sampling_loc / attn_weight must be float32")
    B, S, M, D = value.shape
    Bq, Nq, Mq, L, P, _ = sampling_locations.shape
    _require(B == Bq and M == Mq, "batch / head mismatch between value and sampling_loc")
    _require(spatial_shapes.shape == (L, 2) and level_start_index.shape == (L,), "spatial_shapes must be [L, 2], level_start_index [L]")
    return B, S, M, D, L, Nq, P


def _grad_like_param(g: Tensor, shape) -> Tensor:
    """``g`` ([H, 64]) viewed with the parameter's shape AND the canonical contiguous strides of that shape: a plain
    ``view`` to [H, 64, 1, 1] keeps stride 64 on the size-1 dims, which DDP's reducer reports as a layout mismatch with its
    bucket view ("grad strides do not match bucket view strides") and handles by an extra copy."""
    shape = tuple(shape)
    strides, acc = [], 1
    for n in reversed(shape):
        strides.append(acc)
        acc *= max(int(n), 1)
    return g.as_strided(shape, tuple(reversed(strides)))


# ---- MSDA ---------------------------------------------------------------------------------------

@torch.library.custom_op("rdetr::msda_forward", mutates_args=(), device_types="cuda")
def msda_forward(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor, sampling_locations: Tensor,
                 attention_weights: Tensor) -> Tensor:
    B, S, M, D, L, Nq, P = _check_msda_inputs(value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
    out = torch.empty((B, Nq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        rc = _lib.lib().rdetr_msda_forward(_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index),
                                           _ptr(sampling_locations), _ptr(attention_weights), _ptr(out),
                                           B, S, M, D, L, Nq, P, _value_dtype_code(value), _stream(value))
    _lib.check(rc, "rdetr_msda_forward")
    return out


@msda_forward.register_fake
def _(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    B, _, M, D = value.shape
    return value.new_empty((B, sampling_locations.shape[1], M * D))


@torch.library.custom_op("rdetr::msda_backward", mutates_args=(), device_types="cuda")
def msda_backward(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor, sampling_locations: Tensor,
                  attention_weights: Tensor, grad_output: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    B, S, M, D, L, Nq, P = _check_msda_inputs(value, spatial_shapes, level_start_index, sampling_locations, attention_weights)
    _require(grad_output.is_cuda, "grad_output must be a CUDA tensor")
    grad_output = grad_output.to(value.dtype).contiguous()  # the reference asserts contiguity (cu:90); we normalise
    _require(grad_output.shape == (B, Nq, M * D), "grad_output must be [B, Nq, M*D]")
    code = _value_dtype_code(value)
    grad_value = torch.empty_like(value)  # zeroed inside the library
    grad_loc = torch.empty_like(sampling_locations)
    grad_attn = torch.empty_like(attention_weights)
    L_ = _lib.lib()
    ws_bytes = L_.rdetr_msda_backward_workspace_bytes(B, S, M, D, L, Nq, P, code)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=value.device) if ws_bytes else None
    with torch.cuda.device(value.device):
        rc = L_.rdetr_msda_backward(_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_locations),
                                    _ptr(attention_weights), _ptr(grad_output), _ptr(grad_value), _ptr(grad_loc),
                                    _ptr(grad_attn), B, S, M, D, L, Nq, P, code, _ptr(ws), ws_bytes, _stream(value))
    _lib.check(rc, "rdetr_msda_backward")
    return grad_value, grad_loc, grad_attn


@msda_backward.register_fake
def _(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, grad_output):
    return torch.empty_like(value), torch.empty_like(sampling_locations), torch.empty_like(attention_weights)


def _msda_setup_context(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _msda_autograd_backward(ctx, grad_output):
    value, spatial_shapes, level_start_index, sampling_locations, attention_weights = ctx.saved_tensors
    gv, gl, ga = msda_backward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, grad_output)
    return gv, None, None, gl, ga


msda_forward.register_autograd(_msda_autograd_backward, setup_context=_msda_setup_context)


def ms_deform_attn(value, value_spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """Functional form; differentiable w.r.t. value, sampling_locations, attention_weights
    (once-differentiable, like the reference: ms_deform_attn.py:64-65)."""
    return msda_forward(value, value_spatial_shapes, level_start_index, sampling_locations, attention_weights)


class MultiScaleDeformableAttnFunction:
    """Call-compatible stand-in for the reference autograd.Function (ms_deform_attn.py:35-84):
    ``MultiScaleDeformableAttnFunction.apply(value, shapes, level_start, loc, attn, im2col_step)``."""

    @staticmethod
    def apply(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step=64):
        batch = value.shape[0]
        step = min(batch, int(im2col_step))
        # the reference rejects batches that are not a multiple of the step (ms_deform_attn_cuda.cu:44)
        _require(batch == 0 or (step > 0 and batch % step == 0), f"batch({batch}) must divide im2col_step({step})")
        return msda_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)


# ---- MSDA with the module prologue fused (softmax, location arithmetic, padding mask) -------------

def _check_fused(value, spatial_shapes, level_start_index, reference_points, offsets, logits, key_padding_mask):
    for name, t in (("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                    ("reference_points", reference_points), ("sampling_offsets", offsets), ("attention_logits", logits)):
        _require(t.is_cuda, f"{name} must be a CUDA tensor")
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
    _require(value.dim() == 4 and offsets.dim() == 6 and offsets.shape[-1] == 2, "value [B,S,M,D], sampling_offsets [B,Nq,M,L,P,2]")
    B, S, M, D = value.shape
    Bq, Nq, Mq, L, P, _ = offsets.shape
    _require(B == Bq and M == Mq, "batch / head mismatch between value and sampling_offsets")
    _require(logits.shape == (B, Nq, M, L * P), "attention_logits must be [B, Nq, M, L*P]")
    _require(reference_points.dtype == torch.float32 and reference_points.dim() == 4
             and reference_points.shape[:3] == (B, Nq, L), "reference_points must be fp32 [B, Nq, L, 2 or 4]")
    R = reference_points.shape[-1]
    if R not in (2, 4):
        raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(R))
    _require(offsets.dtype == value.dtype and logits.dtype == value.dtype, "sampling_offsets / attention_logits must have value's dtype")
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64
             and spatial_shapes.shape == (L, 2) and level_start_index.shape == (L,), "spatial_shapes [L,2] / level_start_index [L] int64")
    if key_padding_mask is not None:
        _require(key_padding_mask.is_cuda and key_padding_mask.dtype == torch.bool and key_padding_mask.is_contiguous()
                 and key_padding_mask.shape == (B, S), "key_padding_mask must be a contiguous CUDA bool tensor [B, S]")
    return B, S, M, D, L, Nq, P, R


@torch.library.custom_op("rdetr::msda_fused_forward", mutates_args=(), device_types="cuda")
def msda_fused_forward(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor, reference_points: Tensor,
                       sampling_offsets: Tensor, attention_logits: Tensor, key_padding_mask: Optional[Tensor]) -> Tensor:
    B, S, M, D, L, Nq, P, R = _check_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                          attention_logits, key_padding_mask)
    out = torch.empty((B, Nq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        rc = _lib.lib().rdetr_msda_fused_forward(_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index),
                                                 _ptr(reference_points), _ptr(sampling_offsets), _ptr(attention_logits),
                                                 _ptr(key_padding_mask), _ptr(out), B, S, M, D, L, Nq, P, R,
                                                 _value_dtype_code(value), _stream(value))
    _lib.check(rc, "rdetr_msda_fused_forward")
    return out


@msda_fused_forward.register_fake
def _(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits, key_padding_mask):
    B, _, M, D = value.shape
    return value.new_empty((B, sampling_offsets.shape[1], M * D))


@torch.library.custom_op("rdetr::msda_fused_backward", mutates_args=(), device_types="cuda")
def msda_fused_backward(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor, reference_points: Tensor,
                        sampling_offsets: Tensor, attention_logits: Tensor, key_padding_mask: Optional[Tensor],
                        grad_output: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    B, S, M, D, L, Nq, P, R = _check_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                          attention_logits, key_padding_mask)
    grad_output = grad_output.to(value.dtype).contiguous()
    _require(grad_output.shape == (B, Nq, M * D), "grad_output must be [B, Nq, M*D]")
    code = _value_dtype_code(value)
    grad_value = torch.empty_like(value)
    grad_offsets = torch.empty_like(sampling_offsets)
    grad_logits = torch.empty_like(attention_logits)
    L_ = _lib.lib()
    ws_bytes = L_.rdetr_msda_backward_workspace_bytes(B, S, M, D, L, Nq, P, code)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=value.device) if ws_bytes else None
    with torch.cuda.device(value.device):
        rc = L_.rdetr_msda_fused_backward(_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(reference_points),
                                          _ptr(sampling_offsets), _ptr(attention_logits), _ptr(key_padding_mask),
                                          _ptr(grad_output), _ptr(grad_value), _ptr(grad_offsets), _ptr(grad_logits),
                                          B, S, M, D, L, Nq, P, R, code, _ptr(ws), ws_bytes, _stream(value))
    _lib.check(rc, "rdetr_msda_fused_backward")
    return grad_value, grad_offsets, grad_logits


@msda_fused_backward.register_fake
def _(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits, key_padding_mask, grad_output):
    return torch.empty_like(value), torch.empty_like(sampling_offsets), torch.empty_like(attention_logits)


def _fused_setup_context(ctx, inputs, output):
    value, spatial_shapes, level_start_index, reference_points, offsets, logits, mask = inputs
    ctx.save_for_backward(value, spatial_shapes, level_start_index, reference_points, offsets, logits)
    ctx.mask = mask  # non-differentiable bool tensor


def _fused_autograd_backward(ctx, grad_output):
    value, spatial_shapes, level_start_index, reference_points, offsets, logits = ctx.saved_tensors
    gv, go, gl = msda_fused_backward(value, spatial_shapes, level_start_index, reference_points, offsets, logits, ctx.mask,
                                     grad_output)
    # reference points are detached / constant at every call site of the reference
    # (relation_transformer.py:337, base_transformer.py:57-75): no gradient is produced for them
    return gv, None, None, None, go, gl, None


msda_fused_forward.register_autograd(_fused_autograd_backward, setup_context=_fused_setup_context)


def ms_deform_attn_fused(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits,
                         key_padding_mask=None):
    """value [B,S,M,D] (unmasked), reference_points [B,Nq,L,2|4] fp32, raw sampling_offsets [B,Nq,M,L,P,2] and
    pre-softmax attention_logits [B,Nq,M,L*P] -> [B,Nq,M*D]; differentiable w.r.t. value, offsets, logits."""
    return msda_fused_forward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                              attention_logits, key_padding_mask)


# ---- REL ----------------------------------------------------------------------------------------

def relation_dim_t(num_pos_feats: int = 16, temperature: float = 10000.0, device=None) -> Tensor:
    """``temperature ** (2k / num_pos_feats)`` in fp32, computed by torch exactly as the reference
    does (position_encoding.py:101-105) so the kernel divides by bit-identical constants."""
    dim_t = torch.arange(num_pos_feats // 2, dtype=torch.float32, device=device)
    return temperature ** (dim_t * 2 / num_pos_feats)


def _check_rel(src_boxes, tgt_boxes, weight, bias, dim_t):
    for name, t in (("src_boxes", src_boxes), ("tgt_boxes", tgt_boxes), ("weight", weight), ("bias", bias), ("dim_t", dim_t)):
        _require(t.is_cuda, f"{name} must be a CUDA tensor")
        _require(t.dtype == torch.float32, f"{name} must be float32")
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
    _require(src_boxes.dim() == 3 and src_boxes.shape[-1] == 4, "src_boxes much have 4 coordinates")
    _require(tgt_boxes.dim() == 3 and tgt_boxes.shape[-1] == 4, "tgt_boxes must have 4 coordinates")
    _require(src_boxes.shape[0] == tgt_boxes.shape[0], "src_boxes / tgt_boxes batch mismatch")
    H = bias.shape[0]
    _require(weight.numel() == H * 64 and dim_t.numel() == 8, "weight must hold [H, 64] values and dim_t 8")
    return src_boxes.shape[0], src_boxes.shape[1], tgt_boxes.shape[1], H


@torch.library.custom_op("rdetr::relation_forward", mutates_args=(), device_types="cuda")
def relation_forward(src_boxes: Tensor, tgt_boxes: Tensor, weight: Tensor, bias: Tensor, dim_t: Tensor, scale: float,
                     eps: float, attn_mask: Optional[Tensor], fast: bool) -> Tuple[Tensor, Tensor]:
    B, N1, N2, H = _check_rel(src_boxes, tgt_boxes, weight, bias, dim_t)
    if attn_mask is not None:
        _require(attn_mask.is_cuda and attn_mask.dtype == torch.bool and attn_mask.is_contiguous()
                 and attn_mask.shape == (N1, N2), "attn_mask must be a contiguous CUDA bool tensor of shape [N1, N2]")
    out = torch.empty((B, H, N1, N2), dtype=torch.float32, device=src_boxes.device)
    bits = torch.empty((B, N1, (N2 + 31) // 32, H), dtype=torch.int32, device=src_boxes.device)
    flags = _lib.REL_FAST if fast else _lib.REL_EXACT
    L_ = _lib.lib()
    ws_bytes = L_.rdetr_relation_workspace_bytes(B, N1, N2, flags)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=src_boxes.device) if ws_bytes else None
    with torch.cuda.device(src_boxes.device):
        rc = L_.rdetr_relation_forward(_ptr(src_boxes), _ptr(tgt_boxes), _ptr(weight), _ptr(bias), _ptr(dim_t),
                                       float(scale), float(eps), _ptr(attn_mask), _ptr(out), _ptr(bits),
                                       B, N1, N2, H, flags, _ptr(ws), ws_bytes, _stream(src_boxes))
    _lib.check(rc, "rdetr_relation_forward")
    return out, bits


@relation_forward.register_fake
def _(src_boxes, tgt_boxes, weight, bias, dim_t, scale, eps, attn_mask, fast):
    B, N1, N2, H = src_boxes.shape[0], src_boxes.shape[1], tgt_boxes.shape[1], bias.shape[0]
    return (src_boxes.new_empty((B, H, N1, N2)),
            torch.empty((B, N1, (N2 + 31) // 32, H), dtype=torch.int32, device=src_boxes.device))


@torch.library.custom_op("rdetr::relation_backward", mutates_args=(), device_types="cuda")
def relation_backward(src_boxes: Tensor, tgt_boxes: Tensor, dim_t: Tensor, scale: float, eps: float, grad_out: Tensor,
                      relu_bits: Tensor, num_heads: int, fast: bool) -> Tuple[Tensor, Tensor]:
    B, N1, N2 = src_boxes.shape[0], src_boxes.shape[1], tgt_boxes.shape[1]
    grad_out = grad_out.to(torch.float32).contiguous()
    _require(grad_out.shape == (B, num_heads, N1, N2), "grad_out must be [B, H, N1, N2]")
    gw = torch.empty((num_heads, 64), dtype=torch.float32, device=src_boxes.device)  # zeroed inside the library
    gb = torch.empty((num_heads,), dtype=torch.float32, device=src_boxes.device)
    flags = _lib.REL_FAST if fast else _lib.REL_EXACT
    L_ = _lib.lib()
    ws_bytes = L_.rdetr_relation_workspace_bytes(B, N1, N2, flags)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=src_boxes.device) if ws_bytes else None
    with torch.cuda.device(src_boxes.device):
        rc = L_.rdetr_relation_backward(_ptr(src_boxes), _ptr(tgt_boxes), _ptr(dim_t), float(scale), float(eps),
                                        _ptr(grad_out), _ptr(relu_bits), _ptr(gw), _ptr(gb), B, N1, N2, num_heads,
                                        flags, _ptr(ws), ws_bytes, _stream(src_boxes))
    _lib.check(rc, "rdetr_relation_backward")
    return gw, gb


@relation_backward.register_fake
def _(src_boxes, tgt_boxes, dim_t, scale, eps, grad_out, relu_bits, num_heads, fast):
    return src_boxes.new_empty((num_heads, 64)), src_boxes.new_empty((num_heads,))


def _rel_setup_context(ctx, inputs, output):
    src_boxes, tgt_boxes, weight, bias, dim_t, scale, eps, attn_mask, fast = inputs
    _, bits = output
    # the output itself is NOT saved: the decoder mutates it in place (relation_transformer.py:372-374),
    # which would trip autograd's version check; the 1-bit ReLU mask is what the backward needs.
    ctx.save_for_backward(src_boxes, tgt_boxes, dim_t, bits)
    ctx.scale, ctx.eps, ctx.fast = scale, eps, fast
    ctx.num_heads = bias.shape[0]
    ctx.weight_shape = weight.shape


def _rel_autograd_backward(ctx, grad_out, grad_bits):
    src_boxes, tgt_boxes, dim_t, bits = ctx.saved_tensors
    gw, gb = relation_backward(src_boxes, tgt_boxes, dim_t, ctx.scale, ctx.eps, grad_out, bits, ctx.num_heads, ctx.fast)
    # no gradient reaches the boxes: the reference computes the geometry under no_grad (:527-529)
    return None, None, _grad_like_param(gw, ctx.weight_shape), gb, None, None, None, None, None


relation_forward.register_autograd(_rel_autograd_backward, setup_context=_rel_setup_context)


def position_relation_bias(src_boxes: Tensor, tgt_boxes: Optional[Tensor], weight: Tensor, bias: Tensor,
                           dim_t: Optional[Tensor] = None, scale: float = 100.0, eps: float = 1e-5,
                           attn_mask: Optional[Tensor] = None, fast: bool = False) -> Tensor:
    """boxes in -> ``[B, H, N1, N2]`` fp32 bias out (>= 0, or -inf where ``attn_mask``); differentiable
    w.r.t. ``weight`` ([H,64] or [H,64,1,1]) and ``bias`` only."""
    if tgt_boxes is None:
        tgt_boxes = src_boxes
    if dim_t is None:
        dim_t = relation_dim_t(16, 10000.0, src_boxes.device)
    out, _ = relation_forward(src_boxes.detach().float().contiguous(), tgt_boxes.detach().float().contiguous(),
                              weight.float().contiguous(), bias.float().contiguous(), dim_t, scale, eps, attn_mask, fast)
    return out


# ---- fused relation attention -----------------------------------------

def _check_relattn(q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, attn_mask):
    for name, t in (("q", q), ("k", k), ("v", v), ("src_boxes", src_boxes), ("tgt_boxes", tgt_boxes), ("weight", weight),
                    ("bias", bias), ("dim_t", dim_t)):
        _require(t.is_cuda, f"{name} must be a CUDA tensor")
        _require(t.dtype == torch.float32, f"{name} must be float32")
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
    _require(q.dim() == 4 and q.shape == k.shape == v.shape, "q, k, v must be [B, H, N, D] of one shape")
    B, H, N, D = q.shape
    _require(src_boxes.shape == (B, N, 4) and tgt_boxes.shape == (B, N, 4), "boxes must be [B, N, 4]")
    _require(weight.numel() == H * 64 and bias.numel() == H and dim_t.numel() == 8, "weight must hold [H, 64] values, bias H, dim_t 8")
    if attn_mask is not None:
        _require(attn_mask.is_cuda and attn_mask.dtype == torch.bool and attn_mask.is_contiguous() and attn_mask.shape == (N, N),
                 "attn_mask must be a contiguous CUDA bool tensor of shape [N, N]")
    return B, H, N, D


@torch.library.custom_op("rdetr::relation_attention_forward", mutates_args=(), device_types="cuda")
def relation_attention_forward(q: Tensor, k: Tensor, v: Tensor, src_boxes: Tensor, tgt_boxes: Tensor, weight: Tensor, bias: Tensor,
                               dim_t: Tensor, scale: float, eps: float, attn_mask: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    B, H, N, D = _check_relattn(q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, attn_mask)
    out = torch.empty_like(q)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=q.device)
    L_ = _lib.lib()
    ws_bytes = L_.rdetr_relation_attention_workspace_bytes(B, N, H, 0)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=q.device) if ws_bytes else None
    with torch.cuda.device(q.device):
        rc = L_.rdetr_relation_attention_forward(_ptr(q), _ptr(k), _ptr(v), _ptr(src_boxes), _ptr(tgt_boxes), _ptr(weight), _ptr(bias),
                                                 _ptr(dim_t), float(scale), float(eps), _ptr(attn_mask), _ptr(out), _ptr(lse),
                                                 B, N, H, D, _ptr(ws), ws_bytes, _stream(q))
    _lib.check(rc, "rdetr_relation_attention_forward")
    return out, lse


@relation_attention_forward.register_fake
def _(q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, scale, eps, attn_mask):
    return torch.empty_like(q), q.new_empty(q.shape[:3])


@torch.library.custom_op("rdetr::relation_attention_backward", mutates_args=(), device_types="cuda")
def relation_attention_backward(q: Tensor, k: Tensor, v: Tensor, src_boxes: Tensor, tgt_boxes: Tensor, weight: Tensor, bias: Tensor,
                                dim_t: Tensor, scale: float, eps: float, attn_mask: Optional[Tensor], out: Tensor, lse: Tensor,
                                grad_out: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    B, H, N, D = _check_relattn(q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, attn_mask)
    grad_out = grad_out.to(torch.float32).contiguous()
    _require(grad_out.shape == q.shape, "grad_out must have the shape of q")
    gq, gk, gv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)   # gk / gv are zeroed inside the library
    gw = torch.empty((H, 64), dtype=torch.float32, device=q.device)
    gb = torch.empty((H,), dtype=torch.float32, device=q.device)
    L_ = _lib.lib()
    ws_bytes = L_.rdetr_relation_attention_workspace_bytes(B, N, H, 1)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=q.device) if ws_bytes else None
    with torch.cuda.device(q.device):
        rc = L_.rdetr_relation_attention_backward(_ptr(q), _ptr(k), _ptr(v), _ptr(src_boxes), _ptr(tgt_boxes), _ptr(weight), _ptr(bias),
                                                  _ptr(dim_t), float(scale), float(eps), _ptr(attn_mask), _ptr(out), _ptr(lse),
                                                  _ptr(grad_out), _ptr(gq), _ptr(gk), _ptr(gv), _ptr(gw), _ptr(gb), B, N, H, D,
                                                  _ptr(ws), ws_bytes, _stream(q))
    _lib.check(rc, "rdetr_relation_attention_backward")
    return gq, gk, gv, gw, gb


@relation_attention_backward.register_fake
def _(q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, scale, eps, attn_mask, out, lse, grad_out):
    H = q.shape[1]
    return torch.empty_like(q), torch.empty_like(q), torch.empty_like(q), q.new_empty((H, 64)), q.new_empty((H,))


def _relattn_setup_context(ctx, inputs, output):
    q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, scale, eps, attn_mask = inputs
    out, lse = output
    ctx.save_for_backward(q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, out, lse)
    ctx.attn_mask = attn_mask
    ctx.scale, ctx.eps = scale, eps
    ctx.weight_shape = weight.shape


def _relattn_autograd_backward(ctx, grad_out, grad_lse):
    q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, out, lse = ctx.saved_tensors
    gq, gk, gv, gw, gb = relation_attention_backward(q, k, v, src_boxes, tgt_boxes, weight, bias, dim_t, ctx.scale, ctx.eps,
                                                     ctx.attn_mask, out, lse, grad_out)
    # no gradient reaches the boxes: the reference computes the geometry under no_grad (relation_transformer.py:527-529)
    return gq, gk, gv, None, None, _grad_like_param(gw, ctx.weight_shape), gb, None, None, None, None


relation_attention_forward.register_autograd(_relattn_autograd_backward, setup_context=_relattn_setup_context)


def relation_attention(q: Tensor, k: Tensor, v: Tensor, src_boxes: Tensor, tgt_boxes: Tensor, weight: Tensor, bias: Tensor,
                       dim_t: Optional[Tensor] = None, scale: float = 100.0, eps: float = 1e-5,
                       attn_mask: Optional[Tensor] = None) -> Tensor:
    """``softmax(q k^T / sqrt(D) + relu(W f(src, tgt) + c) [-inf where attn_mask]) v`` for ``q, k, v [B, H, N, D]`` without ever
    materialising the ``[B, H, N, N]`` relation bias (reference: relation_transformer.py:369-374 + :453-459).  Differentiable
    w.r.t. q, k, v, ``weight`` ([H,64] or [H,64,1,1]) and ``bias``."""
    if dim_t is None:
        dim_t = relation_dim_t(16, 10000.0, q.device)
    out, _ = relation_attention_forward(q.float().contiguous(), k.float().contiguous(), v.float().contiguous(),
                                        src_boxes.detach().float().contiguous(), tgt_boxes.detach().float().contiguous(),
                                        weight.float().contiguous(), bias.float().contiguous(), dim_t, scale, eps, attn_mask)
    return out


# ---- memory_fusion input Linear without the concatenation --------------

@torch.library.custom_op("rdetr::memory_fusion_forward", mutates_args=(), device_types="cuda")
def memory_fusion_forward(sources: Sequence[Tensor], weight: Tensor, bias: Tensor, relu: bool) -> Tensor:
    import ctypes

    srcs = list(sources)
    _require(len(srcs) > 0, "rdetr::memory_fusion: no source tensor")
    C = srcs[0].shape[-1]
    lead = srcs[0].shape[:-1]
    flat = []
    for t in srcs:
        _require(t.is_cuda and t.dtype == torch.float32, "rdetr::memory_fusion: sources must be float32 CUDA tensors (no CPU path)")
        _require(t.shape[-1] == C and t.shape[:-1] == lead, "rdetr::memory_fusion: sources must share one shape")
        flat.append(t.reshape(-1, C).contiguous())
    N = weight.shape[0]
    _require(weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous() and weight.shape == (N, len(srcs) * C),
             "rdetr::memory_fusion: weight must be a contiguous float32 [N, len(sources)*C] CUDA tensor")
    _require(bias.is_cuda and bias.dtype == torch.float32 and bias.is_contiguous() and bias.shape == (N,), "rdetr::memory_fusion: bias must be float32 [N]")
    M = flat[0].shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=weight.device)
    ptrs = (ctypes.c_void_p * len(flat))(*[t.data_ptr() for t in flat])
    with torch.cuda.device(weight.device):
        rc = _lib.lib().rdetr_memory_fusion_forward(ptrs, len(flat), _ptr(weight), _ptr(bias), _ptr(out), M, C, N, 1 if relu else 0,
                                                    _stream(weight))
    _lib.check(rc, "rdetr_memory_fusion_forward")
    return out.view(*lead, N)


@memory_fusion_forward.register_fake
def _(sources, weight, bias, relu):
    return sources[0].new_empty((*sources[0].shape[:-1], weight.shape[0]))


def _memfuse_setup_context(ctx, inputs, output):
    sources, weight, bias, relu = inputs
    ctx.save_for_backward(weight, output, *sources)  # the sources themselves, not a concatenated copy
    ctx.relu = relu
    ctx.autocast = torch.is_autocast_enabled()


def _memfuse_autograd_backward(ctx, grad_out):
    """Plain library GEMMs on the original tensors: dX_t = g W_t, dW_t = g^T X_t, at the precision the forward ran the contraction
    in -- bf16 when the forward was called under autocast (what upstream's autocast backward does), TF32 otherwise."""
    weight, output, *sources = ctx.saved_tensors
    C, N = sources[0].shape[-1], weight.shape[0]
    g = grad_out.reshape(-1, N).float()
    if ctx.relu:
        g = g * (output.reshape(-1, N) > 0)
    gb = g.sum(0)
    if ctx.autocast:
        g16, w16 = g.bfloat16(), weight.bfloat16()
        grads = [(g16 @ w16[:, t * C:(t + 1) * C]).float().view_as(x) for t, x in enumerate(sources)]
        gw = torch.cat([(g16.t() @ x.reshape(-1, C).bfloat16()).float() for x in sources], 1)
        return grads, gw, gb, None
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        grads = [(g @ weight[:, t * C:(t + 1) * C]).view_as(x) for t, x in enumerate(sources)]
        gw = torch.cat([g.t() @ x.reshape(-1, C) for x in sources], 1)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return grads, gw, gb, None


memory_fusion_forward.register_autograd(_memfuse_autograd_backward, setup_context=_memfuse_setup_context)


def memory_fusion_linear(sources: Sequence[Tensor], weight: Tensor, bias: Tensor, relu: bool = True) -> Tensor:
    """``relu(torch.cat(sources, -1) @ weight.T + bias)`` without the concatenation: a K-split tcgen05 GEMM (TF32 products, fp32
    accumulation) whose A operand walks ``sources`` in place (reference: relation_transformer.py:168-173, 203-204)."""
    return memory_fusion_forward([s.float() for s in sources], weight.float().contiguous(), bias.float().contiguous(), relu)


# ---- two-stage query selection (second half) ---------------------------

def topk_rows(scores: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """``torch.topk(scores, k, dim=1)`` for float32 ``scores [B, S]`` on the GPU: one CTA per row (radix select + sort in
    shared memory).  Same values and indices as torch; equal scores come out in ascending index order.  The values are
    gathered with torch so that they stay differentiable, as ``torch.topk``'s are."""
    _require(scores.is_cuda and scores.dtype == torch.float32 and scores.dim() == 2,
             "rdetr::topk_rows: scores must be a float32 [B, S] CUDA tensor (no CPU path)")
    s = scores.detach().contiguous()
    B, S = s.shape
    idx = torch.empty((B, k), dtype=torch.int64, device=s.device)
    with torch.cuda.device(s.device):
        rc = _lib.lib().rdetr_topk_rows(_ptr(s), B, S, int(k), _ptr(idx), 0, _stream(s))
    _lib.check(rc, "rdetr_topk_rows")
    return scores.gather(1, idx), idx


@torch.library.custom_op("rdetr::two_stage_select_forward", mutates_args=(), device_types="cuda")
def two_stage_select_forward(class_logits: Tensor, coord: Tensor, k: int, sigmoid: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _require(class_logits.is_cuda and coord.is_cuda, "rdetr::two_stage_select: CUDA tensors only (no CPU path)")
    _require(class_logits.dtype == torch.float32 and coord.dtype == torch.float32, "rdetr::two_stage_select: float32 inputs")
    _require(class_logits.dim() == 3 and coord.dim() == 3 and coord.shape == (*class_logits.shape[:2], 4),
             "rdetr::two_stage_select: class_logits [B, S, C] and coord [B, S, 4]")
    cls, box = class_logits.contiguous(), coord.contiguous()
    B, S, C = cls.shape
    top_cls = torch.empty((B, k, C), dtype=torch.float32, device=cls.device)
    top_box = torch.empty((B, k, 4), dtype=torch.float32, device=cls.device)
    idx = torch.empty((B, k), dtype=torch.int64, device=cls.device)
    L = _lib.lib()
    nbytes = L.rdetr_two_stage_workspace_bytes(B, S)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=cls.device)
    with torch.cuda.device(cls.device):
        rc = L.rdetr_two_stage_select(_ptr(cls), _ptr(box), B, S, C, int(k), 1 if sigmoid else 0, _ptr(top_cls), _ptr(top_box),
                                      _ptr(idx), _ptr(ws), nbytes, _stream(cls))
    _lib.check(rc, "rdetr_two_stage_select")
    return top_cls, top_box, idx


@two_stage_select_forward.register_fake
def _(class_logits, coord, k, sigmoid):
    B, S, C = class_logits.shape
    return (class_logits.new_empty((B, k, C)), coord.new_empty((B, k, 4)),
            torch.empty((B, k), dtype=torch.int64, device=class_logits.device))


@torch.library.custom_op("rdetr::two_stage_select_backward", mutates_args=(), device_types="cuda")
def two_stage_select_backward(grad_class: Tensor, grad_coord: Tensor, topk_coord: Tensor, topk_index: Tensor, S: int,
                              sigmoid: bool) -> Tuple[Tensor, Tensor]:
    B, K, C = grad_class.shape
    gc, gb = grad_class.float().contiguous(), grad_coord.float().contiguous()
    out_c = torch.empty((B, S, C), dtype=torch.float32, device=gc.device)
    out_b = torch.empty((B, S, 4), dtype=torch.float32, device=gc.device)
    with torch.cuda.device(gc.device):
        rc = _lib.lib().rdetr_two_stage_select_backward(_ptr(gc), _ptr(gb), _ptr(topk_coord), _ptr(topk_index), B, S, C, K,
                                                        1 if sigmoid else 0, _ptr(out_c), _ptr(out_b), _stream(gc))
    _lib.check(rc, "rdetr_two_stage_select_backward")
    return out_c, out_b


@two_stage_select_backward.register_fake
def _(grad_class, grad_coord, topk_coord, topk_index, S, sigmoid):
    B, K, C = grad_class.shape
    return grad_class.new_empty((B, S, C)), grad_coord.new_empty((B, S, 4))


def _select_setup_context(ctx, inputs, output):
    class_logits, coord, k, sigmoid = inputs
    ctx.save_for_backward(output[1], output[2])
    ctx.S, ctx.sigmoid = class_logits.shape[1], sigmoid
    ctx.set_materialize_grads(True)


def _select_autograd_backward(ctx, grad_class, grad_coord, grad_index):
    topk_coord, topk_index = ctx.saved_tensors
    gc, gb = two_stage_select_backward(grad_class, grad_coord, topk_coord, topk_index, ctx.S, ctx.sigmoid)
    return gc, gb, None, None


two_stage_select_forward.register_autograd(_select_autograd_backward, setup_context=_select_setup_context)


def two_stage_select(class_logits: Tensor, coord_unact: Tensor, k: int, sigmoid: bool = True) -> Tuple[Tensor, Tensor, Tensor]:
    """The reference's two-stage query selection (``relation_transformer.py:90-96`` / ``:104-111``) in three launches:
    ``(class_logits.gather(topk rows), coord_unact.sigmoid().gather(topk rows), topk_index)`` with
    ``topk_index = torch.topk(class_logits.max(-1)[0], k, dim=1)[1]``; differentiable in both inputs."""
    top_cls, top_box, idx = two_stage_select_forward(class_logits.float(), coord_unact.float(), int(k), bool(sigmoid))
    return top_cls.to(class_logits.dtype), top_box.to(coord_unact.dtype), idx


# ---- batched linear-sum-assignment -------------------------------------

def lsap_solve(costs: Sequence[Tensor]) -> Tuple[List[Tuple[Tensor, Tensor]], Tensor]:
    """Solve every cost matrix of ``costs`` (float32 CUDA tensors ``[n_rows, n_cols]``, same device) with one
    launch of the device solver.  -> (``[(row_ind, col_ind), ...]`` int64 device tensors with the pairs
    ``scipy.optimize.linear_sum_assignment`` returns for the same matrix, ``status`` int32 ``[len(costs)]``
    on the device: 0 ok, 1 infeasible, 2 NaN/-inf entry).  Replaces ``linear_sum_assignment(c.cpu())``
    (reference models/matcher/hungarian_matcher.py:80,87) without a host synchronisation: the status is
    left on the device for the caller to inspect when it chooses to."""
    import ctypes

    costs = list(costs)
    _require(len(costs) > 0, "rdetr::lsap: no cost matrix")
    dev = costs[0].device
    mats = []
    for c in costs:
        _require(c.is_cuda and c.device == dev, "rdetr::lsap: cost matrices must be CUDA tensors on one device (no CPU path)")
        _require(c.dim() == 2 and c.dtype == torch.float32, f"rdetr::lsap: expected float32 [rows, cols], got {c.dtype} {tuple(c.shape)}")
        mats.append(c.detach().contiguous())
    P = len(mats)
    lens = [min(c.shape) for c in mats]
    flat = torch.empty(2 * sum(lens), dtype=torch.int64, device=dev)
    status = torch.empty(P, dtype=torch.int32, device=dev)
    out, cursor = [], 0
    for n in lens:
        out.append((flat[cursor:cursor + n], flat[cursor + n:cursor + 2 * n]))
        cursor += 2 * n
    n_rows = (ctypes.c_int64 * P)(*[c.shape[0] for c in mats])
    n_cols = (ctypes.c_int64 * P)(*[c.shape[1] for c in mats])
    cost_p = (ctypes.c_void_p * P)(*[c.data_ptr() for c in mats])
    row_p = (ctypes.c_void_p * P)(*[r.data_ptr() for r, _ in out])
    col_p = (ctypes.c_void_p * P)(*[c.data_ptr() for _, c in out])
    lib = _lib.lib()
    ws_bytes = lib.rdetr_lsap_workspace_bytes(n_rows, n_cols, P)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    _lib.check(lib.rdetr_lsap_solve(cost_p, n_rows, n_cols, row_p, col_p, _ptr(status), P, _ptr(ws), ws_bytes, _stream(status)),
               "rdetr_lsap_solve")
    return out, status


def match_cost(pred_boxes: Sequence[Tensor], pred_logits: Sequence[Tensor], gt_boxes: Sequence[Tensor],
               gt_labels: Sequence[Tensor], cost_class: float, cost_bbox: float, cost_giou: float,
               focal_alpha: float = 0.25, focal_gamma: float = 2.0) -> List[Tensor]:
    """Cost matrices ``[n_queries, n_gt]`` of a batch of matching problems in one launch: the values (bit for
    bit) that ``HungarianMatcher.calculate_cost`` of the reference computes with ~30 eager kernels per problem
    (models/matcher/hungarian_matcher.py:40-72)."""
    import ctypes

    P = len(pred_boxes)
    _require(P > 0 and len(pred_logits) == P and len(gt_boxes) == P and len(gt_labels) == P, "rdetr::match_cost: ragged argument lists")
    dev = pred_boxes[0].device
    num_classes = pred_logits[0].shape[-1]
    keep, nq, ng = [], [], []
    for pb, pl, gb, gl in zip(pred_boxes, pred_logits, gt_boxes, gt_labels):
        for t in (pb, pl, gb, gl):
            _require(t.is_cuda and t.device == dev, "rdetr::match_cost: tensors must be CUDA tensors on one device (no CPU path)")
        _require(pb.dtype == torch.float32 and pl.dtype == torch.float32 and gb.dtype == torch.float32 and gl.dtype == torch.int64,
                 "rdetr::match_cost: float32 boxes / logits and int64 labels expected")
        _require(pb.dim() == 2 and pb.shape[1] == 4 and gb.dim() == 2 and gb.shape[1] == 4 and pl.dim() == 2
                 and pl.shape[0] == pb.shape[0] and pl.shape[1] == num_classes and gl.shape == gb.shape[:1],
                 f"rdetr::match_cost: shapes {tuple(pb.shape)} {tuple(pl.shape)} {tuple(gb.shape)} {tuple(gl.shape)}")
        pb, gb = pb.detach().contiguous(), gb.detach().contiguous()
        pb = pb if pb.data_ptr() % 16 == 0 else pb.clone()          # float4 loads
        gb = gb if gb.data_ptr() % 16 == 0 else gb.clone()
        keep.append((pb, pl.detach().contiguous(), gb, gl.contiguous()))
        nq.append(pb.shape[0])
        ng.append(gb.shape[0])
    # one allocation for all matrices; every matrix starts on a 16-byte boundary
    starts, cursor = [], 0
    for a, b in zip(nq, ng):
        starts.append(cursor)
        cursor += (a * b + 3) // 4 * 4
    flat = torch.empty(cursor, dtype=torch.float32, device=dev)
    costs = [flat[s:s + a * b].view(a, b) for s, a, b in zip(starts, nq, ng)]
    arr = lambda xs: (ctypes.c_void_p * P)(*xs)  # noqa: E731
    _lib.check(_lib.lib().rdetr_match_cost(
        arr([k[0].data_ptr() for k in keep]), arr([k[1].data_ptr() for k in keep]), arr([k[2].data_ptr() for k in keep]),
        arr([k[3].data_ptr() for k in keep]), arr([c.data_ptr() for c in costs]), (ctypes.c_int64 * P)(*nq), (ctypes.c_int64 * P)(*ng),
        num_classes, float(cost_class), float(cost_bbox), float(cost_giou), float(focal_alpha), float(focal_gamma), P, _stream(flat)),
        "rdetr_match_cost")
    return costs
