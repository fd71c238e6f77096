"""Named workloads (BASELINE.json ``configs``) and the seeded synthetic-input generators.

The recipes follow SURVEY.md Appendix B; each cites the reference code whose behaviour it mimics.
Everything is generated with an explicit ``torch.Generator`` so the same seed gives the same
tensors on every run of the same device type.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Tuple

import torch


@dataclass(frozen=True)
class MsdaShape:
    """Shape of one MSDA call.  ``levels`` = [(H_l, W_l)] as ``spatial_shapes`` holds them."""
    name: str
    batch: int
    levels: Tuple[Tuple[int, int], ...]
    num_query: int  # 0 => Nq = S (encoder self-attention)
    heads: int = 8
    head_dim: int = 32
    points: int = 4

    @property
    def S(self) -> int:
        return sum(h * w for h, w in self.levels)

    @property
    def Nq(self) -> int:
        return self.num_query or self.S

    @property
    def L(self) -> int:
        return len(self.levels)

    def algorithmic_bytes(self, value_bytes: int = 4) -> Tuple[int, int]:
        """(fwd, bwd) bytes, every tensor counted once (SURVEY.md 8d / BASELINE.md 3).
        value/out/grad_out/grad_value use ``value_bytes`` per element; loc/attn and their grads fp32."""
        B, S, Nq, M, D, L, P = self.batch, self.S, self.Nq, self.heads, self.head_dim, self.L, self.points
        value = B * S * M * D * value_bytes
        out = B * Nq * M * D * value_bytes
        loc = B * Nq * M * L * P * 2 * 4
        attn = B * Nq * M * L * P * 4
        fwd = value + loc + attn + out
        bwd = (out + value + loc + attn) + (value + loc + attn)
        return fwd, bwd


# padded 800x1344 (strides 8..64) and 1216x2016 (strides 4..64, focalnet config) pyramids
LEVELS_800_1333 = ((100, 168), (50, 84), (25, 42), (13, 21))
LEVELS_1200_2000 = ((304, 504), (152, 252), (76, 126), (38, 63), (19, 32))

MSDA_SHAPES = {
    # configs[1]: the shape the headline metric is quoted on
    "msda_enc_800x1333_b8": MsdaShape("msda_enc_800x1333_b8", 8, LEVELS_800_1333, 0),
    "msda_enc_800x1333_b2": MsdaShape("msda_enc_800x1333_b2", 2, LEVELS_800_1333, 0),
    "msda_enc_800x1333_b1": MsdaShape("msda_enc_800x1333_b1", 1, LEVELS_800_1333, 0),
    "msda_dec_900_b8": MsdaShape("msda_dec_900_b8", 8, LEVELS_800_1333, 900),
    "msda_dec_1100_b8": MsdaShape("msda_dec_1100_b8", 8, LEVELS_800_1333, 1100),
    "msda_dec_1500_b8": MsdaShape("msda_dec_1500_b8", 8, LEVELS_800_1333, 1500),
    "msda_enc_1200x2000_b1": MsdaShape("msda_enc_1200x2000_b1", 1, LEVELS_1200_2000, 0),
    "msda_tiny": MsdaShape("msda_tiny", 2, ((8, 12), (4, 6), (2, 3)), 37),
}


def shape_tensors(levels, device="cpu"):
    """(spatial_shapes [L,2] int64 (h,w), level_start_index [L] int64) as
    ``models/bricks/base_transformer.py:25-39`` builds them."""
    ss = torch.tensor(list(levels), dtype=torch.int64, device=device)
    lsi = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
    return ss, lsi


def _gen(seed: int, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def grid_init(heads: int, levels: int, points: int) -> torch.Tensor:
    """The ``sampling_offsets.bias`` the module starts from (ms_deform_attn.py:266-278): [M,L,P,2] px."""
    thetas = torch.arange(heads, dtype=torch.float32) * (2.0 * math.pi / heads)
    g = torch.stack([thetas.cos(), thetas.sin()], -1)
    g = g / g.abs().max(-1, keepdim=True)[0]
    g = g.view(heads, 1, 1, 2).repeat(1, levels, points, 1)
    for i in range(points):
        g[:, :, i, :] *= i + 1
    return g


def full_reference_points(levels, device="cpu") -> torch.Tensor:
    """Pixel-centre grid of every level, normalised (base_transformer.py:57-70, valid_ratio = 1): [S,2] (x,y)."""
    pts = []
    for h, w in levels:
        ys, xs = torch.meshgrid(torch.arange(0.5, h + 0.5, device=device),
                                torch.arange(0.5, w + 0.5, device=device), indexing="ij")
        pts.append(torch.stack((xs.reshape(-1) / w, ys.reshape(-1) / h), -1))
    return torch.cat(pts, 0)


def make_loc(shape: MsdaShape, kind: str, seed: int = 0, device="cpu", dtype=torch.float32) -> torch.Tensor:
    """sampling_locations [B,Nq,M,L,P,2].  kind in
    U (uniform, no locality), S (encoder-realistic), D (decoder-realistic), strict (away from pixel
    boundaries, for the grad_loc check), oob (range [-0.15,1.15] so border / outside samples occur)."""
    B, Nq, M, L, P = shape.batch, shape.Nq, shape.heads, shape.L, shape.points
    g = _gen(seed, device)
    wh = torch.tensor([[w, h] for h, w in shape.levels], dtype=torch.float32, device=device)  # (W,H) per level
    if kind == "U":
        loc = torch.rand((B, Nq, M, L, P, 2), generator=g, device=device)
    elif kind == "oob":
        loc = torch.rand((B, Nq, M, L, P, 2), generator=g, device=device) * 1.3 - 0.15
    elif kind == "S":
        ref = full_reference_points(shape.levels, device)
        if Nq != ref.shape[0]:  # decoder-sized call with encoder-like geometry: strided subset of tokens
            idx = torch.linspace(0, ref.shape[0] - 1, Nq, device=device).long()
            ref = ref[idx]
        off = grid_init(M, L, P).to(device)[None, None] + torch.randn((B, Nq, M, L, P, 2), generator=g, device=device)
        loc = ref[None, :, None, None, None, :] + off / wh[None, None, None, :, None, :]
    elif kind == "D":
        cxcy = torch.rand((B, Nq, 2), generator=g, device=device)
        bwh = torch.rand((B, Nq, 2), generator=g, device=device) * 0.48 + 0.02
        off = grid_init(M, L, P).to(device)[None, None] + torch.randn((B, Nq, M, L, P, 2), generator=g, device=device)
        loc = cxcy[:, :, None, None, None, :] + off / P * bwh[:, :, None, None, None, :] * 0.5
    elif kind == "strict":
        # pixel coordinate u = k + f, f in [0.01,0.99] -> loc = (u + 0.5)/size; verified in fp32 below
        size = wh[None, None, None, :, None, :].expand(B, Nq, M, L, P, 2)
        k = torch.floor(torch.rand((B, Nq, M, L, P, 2), generator=g, device=device, dtype=torch.float64)
                        * (size.double() - 1).clamp(min=1))
        f = torch.rand((B, Nq, M, L, P, 2), generator=g, device=device, dtype=torch.float64) * 0.98 + 0.01
        loc = ((k + f + 0.5) / size.double()).float()
        px = loc * size - 0.5
        frac = px - torch.floor(px)
        bad = (frac < 0.005) | (frac > 0.995)
        # offenders are moved to the pixel centre of the same pixel (frac = 0.5)
        loc = torch.where(bad, ((k + 0.5 + 0.5) / size.double()).float(), loc)
    else:
        raise ValueError(f"unknown loc kind {kind!r}")
    return loc.to(dtype).contiguous()


def make_msda_inputs(shape: MsdaShape, loc_kind: str = "U", seed: int = 0, device="cpu", dtype=torch.float32):
    """-> dict(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, grad_output)."""
    B, S, Nq, M, D, L, P = shape.batch, shape.S, shape.Nq, shape.heads, shape.head_dim, shape.L, shape.points
    g = _gen(seed + 1000, device)
    value = torch.randn((B, S, M, D), generator=g, device=device, dtype=torch.float32)
    attn = torch.randn((B, Nq, M, L * P), generator=g, device=device, dtype=torch.float32).softmax(-1)
    grad_out = torch.randn((B, Nq, M * D), generator=g, device=device, dtype=torch.float32)
    ss, lsi = shape_tensors(shape.levels, device)
    return dict(
        value=value.to(dtype).contiguous(),
        spatial_shapes=ss,
        level_start_index=lsi,
        sampling_locations=make_loc(shape, loc_kind, seed, device, dtype),
        attention_weights=attn.view(B, Nq, M, L, P).to(dtype).contiguous(),
        grad_output=grad_out.to(dtype).contiguous(),
    )


@dataclass(frozen=True)
class RelShape:
    name: str
    batch: int
    n1: int
    n2: int
    heads: int = 8

    def algorithmic_bytes(self) -> Tuple[int, int]:
        """(fwd, bwd): the [B,H,N1,N2] fp32 bias written / its gradient read."""
        n = self.batch * self.heads * self.n1 * self.n2 * 4
        return n, n


REL_SHAPES = {
    "rel_900_b8": RelShape("rel_900_b8", 8, 900, 900),
    "rel_1100_b8": RelShape("rel_1100_b8", 8, 1100, 1100),
    "rel_2900_b1": RelShape("rel_2900_b1", 1, 2900, 2900),
    "rel_tiny": RelShape("rel_tiny", 2, 37, 29),
}


def make_boxes(batch: int, n: int, seed: int, device="cpu", dtype=torch.float32, padded_rows: int = 0):
    """cxcy ~ U[0,1), wh ~ U[1e-3,0.5); optional trailing rows = (0.5,0.5,0.5,0.5) (sigmoid(0) slots,
    denoising.py:287-288)."""
    g = _gen(seed, device)
    cxcy = torch.rand((batch, n, 2), generator=g, device=device)
    wh = torch.rand((batch, n, 2), generator=g, device=device) * (0.5 - 1e-3) + 1e-3
    boxes = torch.cat([cxcy, wh], -1)
    if padded_rows:
        boxes[:, n - padded_rows:] = 0.5
    return boxes.to(dtype).contiguous()


def make_rel_params(heads: int = 8, feat: int = 64, seed: int = 0, device="cpu", dtype=torch.float32):
    """Default ``nn.Conv2d(64, heads, 1)`` init (kaiming_uniform a=sqrt(5) => U(+-1/sqrt(fan_in)) = U(+-0.125))."""
    g = _gen(seed + 77, "cpu")
    bound = 1.0 / math.sqrt(feat)
    w = (torch.rand((heads, feat), generator=g) * 2 - 1) * bound
    b = (torch.rand((heads,), generator=g) * 2 - 1) * bound
    return w.to(device=device, dtype=dtype), b.to(device=device, dtype=dtype)


def make_rel_inputs(shape: RelShape, seed: int = 0, device="cpu", dtype=torch.float32):
    src = make_boxes(shape.batch, shape.n1, seed, device, dtype)
    tgt = make_boxes(shape.batch, shape.n2, seed + 1, device, dtype)
    w, b = make_rel_params(shape.heads, 64, seed, device, dtype)
    g = _gen(seed + 2, device)
    grad_out = torch.randn((shape.batch, shape.heads, shape.n1, shape.n2), generator=g, device=device).to(dtype)
    return dict(src_boxes=src, tgt_boxes=tgt, weight=w, bias=b, grad_output=grad_out)


def cdn_attn_mask(num_queries: int, max_gt: int, groups: int, device="cpu") -> torch.Tensor:
    """Bool [N,N] mask the denoising generator builds (denoising.py:66-78): True = blocked."""
    dn = max_gt * groups
    n = dn + num_queries
    mask = torch.zeros((n, n), dtype=torch.bool, device=device)
    mask[dn:, :dn] = True
    for i in range(groups):
        lo, hi = max_gt * i, max_gt * (i + 1)
        mask[lo:hi, :lo] = True
        mask[lo:hi, hi:dn] = True
    return mask
