"""How many grad_value reductions would warp-level merging across x-adjacent queries of one head remove?  (CPU, numpy.)

A warp that held G consecutive queries of the same head could add up contributions that land on the same grad_value row
before sending one reduction (VERDICT r1 weak #3a).  This counts, for configs[1]'s pyramid and sampling locations (one
image), the reductions left under two merging rules:
  slot  : only the same (level, point, corner) slot of the G queries is compared (one shuffle-compare per corner);
  level : all corners of all points of the level within the group are deduplicated (upper bound for any in-warp scheme).
loc kinds: S (module at init: grid offsets + N(0,1) px noise), S0 (the same without the noise: smooth, trained-like).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relation_detr_b200 import workloads  # noqa: E402

shape = workloads.MSDA_SHAPES["msda_enc_800x1333_b1"]
levels = shape.levels
L, P, M = shape.L, shape.points, shape.heads
H0, W0 = levels[0]
nq0 = H0 * W0   # level-0 queries, raster order: x-adjacent queries are consecutive


def corner_rows(loc):
    """loc [Nq, M, L, P, 2] -> rows [Nq, M, L, P, 4] (int64, -1 invalid), per-level local index, as make_tap computes them."""
    out = np.full(loc.shape[:-1] + (4,), -1, dtype=np.int64)
    for l, (H, W) in enumerate(levels):
        w_im = loc[:, :, l, :, 0] * W - 0.5
        h_im = loc[:, :, l, :, 1] * H - 0.5
        inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h0 = np.floor(h_im).astype(np.int64)
        w0 = np.floor(w_im).astype(np.int64)
        for i, (dh, dw) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            hh, ww = h0 + dh, w0 + dw
            ok = inside & (hh >= 0) & (hh <= H - 1) & (ww >= 0) & (ww <= W - 1)
            out[:, :, l, :, i] = np.where(ok, hh * W + ww, -1)
    return out


def report(kind, loc):
    rows = corner_rows(loc)[:nq0]                      # level-0 queries only (75 % of all queries), [nq0, M, L, P, 4]
    total = (rows >= 0).sum(axis=(0, 1, 3, 4)).astype(np.float64)   # per level
    print(f"loc {kind}: reductions per level (level-0 queries, one image): {total.astype(int).tolist()}")
    for G in (2, 4, 8):
        usable = (W0 // G) * G
        r = rows.reshape(H0, W0, M, L, P, 4)[:, :usable].reshape(H0, usable // G, G, M, L, P, 4)
        # slot rule: unique values among the G queries for each (head, level, point, corner)
        s = np.sort(r, axis=2)
        uniq_slot = ((s[:, :, 1:] != s[:, :, :-1]) & (s[:, :, 1:] >= 0)).sum(axis=2) + (s[:, :, 0] >= 0)
        left_slot = uniq_slot.sum(axis=(0, 1, 2, 4, 5)).astype(np.float64)
        # level rule: unique rows among all G * P * 4 corners of the group for each (head, level)
        g = np.moveaxis(r, 2, 4).reshape(H0, usable // G, M, L, G * P * 4)
        s2 = np.sort(g, axis=-1)
        uniq_lvl = ((s2[..., 1:] != s2[..., :-1]) & (s2[..., 1:] >= 0)).sum(axis=-1) + (s2[..., 0] >= 0)
        left_lvl = uniq_lvl.sum(axis=(0, 1, 2)).astype(np.float64)
        tot = (r >= 0).sum(axis=(0, 1, 2, 3, 5, 6)).astype(np.float64)
        print(f"  G = {G}: left after slot merging  per level {np.round(left_slot / tot, 3).tolist()}  all levels {left_slot.sum() / tot.sum():.3f}")
        print(f"         left after level merging per level {np.round(left_lvl / tot, 3).tolist()}  all levels {left_lvl.sum() / tot.sum():.3f}")


loc_s = workloads.make_loc(shape, "S", seed=0)[0].numpy().astype(np.float64)
report("S  (grid offsets + N(0,1) px noise)", loc_s)
# the same without the noise: reference points + grid_init offsets, as smooth as a trained model's offsets can be
ref = workloads.full_reference_points(levels).numpy().astype(np.float64)
wh = np.array([[w, h] for h, w in levels], dtype=np.float64)
off = workloads.grid_init(M, L, P).numpy().astype(np.float64)
loc_s0 = ref[:, None, None, None, :] + off[None] / wh[None, None, :, None, :]
report("S0 (grid offsets only, no noise)", loc_s0)
