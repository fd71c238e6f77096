"""Drop-in ``nn.Module`` replacements for the two reference modules on the hot path.

Constructor arguments, parameter names (``sampling_offsets``, ``attention_weights``, ``value_proj``,
``output_proj``, ``pos_proj.0``) and ``forward`` signatures are those of the reference, so released
checkpoints load with ``strict=True`` and ``optimizer/param_dict.py`` keeps matching by substring:

* ``MultiScaleDeformableAttention``  -- reference ``models/bricks/ms_deform_attn.py:215-377``
* ``PositionRelationEmbedding``      -- reference ``models/bricks/relation_transformer.py:493-532``

Only the dispatch changes: the sampling / relation arithmetic goes to the sm_100a kernels through
``ops.py``; there is no grid_sample or eager fallback.
"""
from __future__ import annotations

import math
import warnings
from typing import Optional

import torch
from torch import Tensor, nn

from . import ops


class MultiScaleDeformableAttention(nn.Module):
    """Multi-scale deformable attention (Deformable-DETR), kernels from ``librdetr_ops.so``."""

    def __init__(self, embed_dim: int = 256, num_levels: int = 4, num_heads: int = 8, num_points: int = 4,
                 img2col_step: int = 64):
        super().__init__()
        if embed_dim % num_heads != 0:
            raise ValueError("embed_dim must be divisible by num_heads, but got {} and {}".format(embed_dim, num_heads))
        head_dim = embed_dim // num_heads
        if head_dim & (head_dim - 1) != 0:
            warnings.warn("embed_dim // num_heads should be a power of 2 for the deformable attention kernels")
        self.im2col_step = img2col_step  # accepted for signature parity; the whole batch is one launch
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.num_levels = num_levels
        self.num_points = num_points
        # True: a bf16 value (autocast) is read as bf16 by the kernel, fp32 accumulation, bf16 out.
        # False: up-cast to fp32 first, exactly as the reference does (ms_deform_attn.py:360).
        self.native_bf16 = True
        # True: softmax, location arithmetic and the padding mask run inside the kernel
        # (rdetr_msda_fused_*); False: computed with torch ops before the call, as the reference does.
        self.fused_prologue = True
        self.sampling_offsets = nn.Linear(embed_dim, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dim, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dim, embed_dim)
        self.output_proj = nn.Linear(embed_dim, embed_dim)
        self.init_weights()

    def init_weights(self):
        """Same initial state as the reference (ms_deform_attn.py:266-284): zero offset weights, a
        ring of per-head directions scaled by the point index as offset bias, uniform attention."""
        nn.init.constant_(self.sampling_offsets.weight.data, 0.0)
        angle = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        ring = torch.stack([angle.cos(), angle.sin()], -1)
        ring = ring / ring.abs().max(-1, keepdim=True)[0]
        ring = ring.view(self.num_heads, 1, 1, 2).repeat(1, self.num_levels, self.num_points, 1)
        ring = ring * torch.arange(1, self.num_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(ring.reshape(-1))
        nn.init.constant_(self.attention_weights.weight.data, 0.0)
        nn.init.constant_(self.attention_weights.bias.data, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight.data)
        nn.init.constant_(self.value_proj.bias.data, 0.0)
        nn.init.xavier_uniform_(self.output_proj.weight.data)
        nn.init.constant_(self.output_proj.bias.data, 0.0)

    def forward(self, query: Tensor, reference_points: Tensor, value: Tensor, spatial_shapes: Tensor,
                level_start_index: Tensor, key_padding_mask: Optional[Tensor]) -> Tensor:
        """query [B,Nq,C]; reference_points [B,Nq,L,2|4] normalised; value [B,S,C]; spatial_shapes [L,2]
        (h,w) int64; level_start_index [L]; key_padding_mask [B,S] bool or None -> [B,Nq,C].

        Unlike the reference there is no ``assert spatial_shapes.prod(1).sum() == S`` here: that
        assert forces a device->host sync on every call (ms_deform_attn.py:313)."""
        B, Nq, _ = query.shape
        S = value.shape[1]
        M, L, P = self.num_heads, self.num_levels, self.num_points

        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
        value = self.value_proj(value)
        in_dtype = value.dtype
        if not (self.native_bf16 and in_dtype == torch.bfloat16):
            value = value.to(torch.float32)
        offsets = self.sampling_offsets(query).view(B, Nq, M, L, P, 2)
        logits = self.attention_weights(query).view(B, Nq, M, L * P)

        if self.fused_prologue:
            if reference_points.requires_grad:
                raise RuntimeError("the fused MSDA prologue produces no gradient for reference_points; detach them "
                                   "(as every call site of the reference does) or set module.fused_prologue = False")
            out = ops.ms_deform_attn_fused(
                value.view(B, S, M, self.embed_dim // M), spatial_shapes, level_start_index,
                reference_points.to(torch.float32).contiguous(), offsets.to(value.dtype).contiguous(),
                logits.to(value.dtype).contiguous(), key_padding_mask)
        else:
            if key_padding_mask is not None:
                value = value.masked_fill(key_padding_mask[..., None], float(0))
            value = value.view(B, S, M, self.embed_dim // M)
            weights = logits.softmax(-1).view(B, Nq, M, L, P)
            if reference_points.shape[-1] == 2:
                wh = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
                locations = reference_points[:, :, None, :, None, :] + offsets / wh[None, None, None, :, None, :]
            else:
                locations = (reference_points[:, :, None, :, None, :2]
                             + offsets / P * reference_points[:, :, None, :, None, 2:] * 0.5)
            out = ops.MultiScaleDeformableAttnFunction.apply(
                value.contiguous(), spatial_shapes, level_start_index,
                locations.to(torch.float32).contiguous(), weights.to(torch.float32).contiguous(), self.im2col_step)
        if out.dtype != in_dtype:
            out = out.to(in_dtype)
        return self.output_proj(out)


class PositionRelationEmbedding(nn.Module):
    """Box-pair geometry -> per-head attention bias ``[B, H, N1, N2]`` in one kernel.

    ``pos_proj`` is kept as ``Sequential(Conv2d(4*embed_dim, H, 1), ReLU)`` purely as the parameter
    container the reference's checkpoints expect (keys ``pos_proj.0.weight`` [H,64,1,1] and
    ``pos_proj.0.bias``); the convolution is never executed."""

    def __init__(self, embed_dim: int = 256, num_heads: int = 8, temperature: float = 10000.0, scale: float = 100.0,
                 activation_layer=nn.ReLU, inplace: bool = True):
        super().__init__()
        if activation_layer is not nn.ReLU:
            raise NotImplementedError("the fused relation kernel implements the ReLU activation of the shipped configs only")
        if embed_dim != 16:
            raise NotImplementedError(f"the fused relation kernel is built for embed_dim=16 (64 features); got {embed_dim}")
        self.pos_proj = nn.Sequential(nn.Conv2d(embed_dim * 4, num_heads, kernel_size=1, bias=True), nn.ReLU(inplace=inplace))
        self.out_channels = num_heads
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.temperature = temperature
        self.scale = scale
        self.eps = 1e-5  # box_rel_encoding default (relation_transformer.py:481)
        # RDETR_REL_FAST by default: measured against the fp64 truth at N=900 it is closer (max 9.1e-6 / mean
        # 3.4e-7) than the reference's own fp32 evaluation (1.3e-5 / 4.5e-7; 4.5e-4 with torch's default
        # TF32 conv) -- profiles/r01_rel_accuracy.json.  False selects the operation-for-operation EXACT mode.
        self.fast_math = True
        self._dim_t = {}

    def _dim_t_on(self, device) -> Tensor:
        key = str(device)
        if key not in self._dim_t:
            self._dim_t[key] = ops.relation_dim_t(self.embed_dim, self.temperature, device)
        return self._dim_t[key]

    #: set by ``install(fused_attention=True)``: ``forward`` then returns a ``LazyRelationBias`` handle that
    #: ``RelationMultiheadAttention`` turns into one fused kernel (row N1); any other consumer materialises it.
    lazy = False

    def forward(self, src_boxes: Tensor, tgt_boxes: Optional[Tensor] = None, attn_mask: Optional[Tensor] = None):
        """src_boxes [B,N1,4], tgt_boxes [B,N2,4] (default: src) -> [B,H,N1,N2], a fresh tensor the
        caller may mutate.  ``attn_mask`` ([N1,N2] bool, optional, not in the reference signature)
        fuses the decoder's ``masked_fill_(attn_mask, -inf)``."""
        if tgt_boxes is None:
            tgt_boxes = src_boxes
        torch._assert(src_boxes.shape[-1] == 4, "src_boxes much have 4 coordinates")
        torch._assert(tgt_boxes.shape[-1] == 4, "tgt_boxes must have 4 coordinates")
        if self.lazy and attn_mask is None and src_boxes.shape[1] == tgt_boxes.shape[1] and self.fast_math:
            return LazyRelationBias(self, src_boxes, tgt_boxes)
        return self._materialize(src_boxes, tgt_boxes, attn_mask)

    def _materialize(self, src_boxes: Tensor, tgt_boxes: Tensor, attn_mask: Optional[Tensor]) -> Tensor:
        conv = self.pos_proj[0]
        out = ops.position_relation_bias(src_boxes, tgt_boxes, conv.weight, conv.bias, self._dim_t_on(src_boxes.device),
                                         self.scale, self.eps, attn_mask, self.fast_math)
        if torch.is_autocast_enabled():
            out = out.to(torch.get_autocast_dtype("cuda"))  # the reference's Conv2d returns the autocast dtype
        return out


class LazyRelationBias:
    """What ``PositionRelationEmbedding`` returns when the fused relation attention is installed: the boxes and the
    embedding's parameters, not the ``[B, H, N, N]`` tensor.  It answers the two calls the reference's decoder makes on the
    tensor (``relation_transformer.py:372-374``: ``.flatten(0, 1)`` and ``.masked_fill_(attn_mask, -inf)``) by recording
    them; ``RelationMultiheadAttention`` consumes it in one kernel, and ``materialize()`` produces the real tensor
    for any other consumer."""

    def __init__(self, embedding: "PositionRelationEmbedding", src_boxes: Tensor, tgt_boxes: Tensor):
        self.embedding = embedding
        self.src_boxes = src_boxes
        self.tgt_boxes = tgt_boxes
        self.mask: Optional[Tensor] = None
        self.flattened = False

    def flatten(self, start_dim: int = 0, end_dim: int = -1):
        if (start_dim, end_dim) != (0, 1):
            return self.materialize().flatten(start_dim, end_dim)
        self.flattened = True
        return self

    def masked_fill_(self, mask: Tensor, value):
        if not (mask.dtype == torch.bool and mask.dim() == 2 and float(value) == float("-inf") and self.mask is None):
            raise NotImplementedError("LazyRelationBias records one masked_fill_(bool [N, N] mask, -inf); materialize() it for anything else")
        self.mask = mask.contiguous()
        return self

    @property
    def shape(self):
        B, N1, N2, H = self.src_boxes.shape[0], self.src_boxes.shape[1], self.tgt_boxes.shape[1], self.embedding.num_heads
        return torch.Size((B * H, N1, N2)) if self.flattened else torch.Size((B, H, N1, N2))

    def materialize(self) -> Tensor:
        out = self.embedding._materialize(self.src_boxes, self.tgt_boxes, self.mask)
        return out.flatten(0, 1) if self.flattened else out


class RelationMultiheadAttention(nn.MultiheadAttention):
    """``nn.MultiheadAttention`` (same parameters and state-dict keys) whose ``attn_mask`` may be a ``LazyRelationBias``:
    the self-attention of ``RelationTransformerDecoderLayer`` (``relation_transformer.py:453-459``) then runs as
    in-projection GEMMs -> ``rdetr::relation_attention_forward`` -> out-projection, and the relation bias is never written
    to memory.  Every other call is ``nn.MultiheadAttention``'s own."""

    def _fusable(self, query, key, value, key_padding_mask, need_weights) -> bool:
        return (self.batch_first and self._qkv_same_embed_dim and self.in_proj_bias is not None and not need_weights
                and key_padding_mask is None and self.bias_k is None and not self.add_zero_attn
                and (self.dropout == 0.0 or not self.training) and query.dim() == 3 and key.shape == query.shape == value.shape
                and self.num_heads == 8 and self.head_dim == 32 and query.is_cuda)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None, average_attn_weights=True,
                is_causal=False):
        if not isinstance(attn_mask, LazyRelationBias):
            return super().forward(query, key, value, key_padding_mask=key_padding_mask, need_weights=need_weights,
                                   attn_mask=attn_mask, average_attn_weights=average_attn_weights, is_causal=is_causal)
        rel = attn_mask
        if not self._fusable(query, key, value, key_padding_mask, need_weights) or rel.src_boxes.shape[1] != query.shape[1]:
            return super().forward(query, key, value, key_padding_mask=key_padding_mask, need_weights=need_weights,
                                   attn_mask=rel.materialize(), average_attn_weights=average_attn_weights, is_causal=is_causal)
        B, N, E = query.shape
        H, D = self.num_heads, self.head_dim
        w, b = self.in_proj_weight, self.in_proj_bias
        if key is query:
            qk = torch.nn.functional.linear(query, w[:2 * E], b[:2 * E])
            q, k = qk[..., :E], qk[..., E:]
        else:
            q = torch.nn.functional.linear(query, w[:E], b[:E])
            k = torch.nn.functional.linear(key, w[E:2 * E], b[E:2 * E])
        v = torch.nn.functional.linear(value, w[2 * E:], b[2 * E:])
        heads = lambda t: t.view(B, N, H, D).transpose(1, 2)  # noqa: E731  ([B, H, N, D]; made contiguous fp32 by the op wrapper)
        emb = rel.embedding
        conv = emb.pos_proj[0]
        core = ops.relation_attention(heads(q), heads(k), heads(v), rel.src_boxes, rel.tgt_boxes, conv.weight, conv.bias,
                                      emb._dim_t_on(query.device), emb.scale, emb.eps, rel.mask)
        core = core.transpose(1, 2).reshape(B, N, E).to(q.dtype)
        return self.out_proj(core), None


def make_fused_encoder(base):
    """``RelationTransformerEncoder`` (upstream ``relation_transformer.py:153-205``) whose last step does not concatenate the
    layer outputs: ``memory_fusion``'s input Linear walks the ``num_layers + 1`` states in place on the tensor cores
    (``rdetr::memory_fusion_forward``).  Same parameters, same state-dict keys; built by
    ``install(fused_memory=True)`` as a subclass of the reference's own class so that the layer loop stays upstream's.

    The kernel multiplies in TF32.  It is used where upstream runs this GEMM at reduced precision anyway -- under autocast
    (bf16 there) or with ``torch.backends.cuda.matmul.allow_tf32`` -- and strict-fp32 calls keep upstream's expression."""

    class RelationTransformerEncoder(base):
        def forward(self, query, spatial_shapes, level_start_index, reference_points, query_pos=None, query_key_padding_mask=None):
            queries = [query]
            for layer in self.layers:
                query = layer(query, query_pos, reference_points, spatial_shapes, level_start_index, query_key_padding_mask)
                queries.append(query)
            lin1, act, lin2, norm = self.memory_fusion
            reduced = torch.is_autocast_enabled() or torch.backends.cuda.matmul.allow_tf32
            if (reduced and query.is_cuda and isinstance(act, nn.ReLU) and lin1.out_features == 256 and len(queries) <= 8
                    and query.shape[-1] % 32 == 0 and lin1.bias is not None):
                hidden = ops.memory_fusion_linear(queries, lin1.weight, lin1.bias, True)
            else:
                hidden = act(lin1(torch.cat(queries, -1)))
            return norm(lin2(hidden))

    RelationTransformerEncoder.__qualname__ = "RelationTransformerEncoder"
    RelationTransformerEncoder.__module__ = base.__module__
    return RelationTransformerEncoder
