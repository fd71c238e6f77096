"""Flat vs tiled MSDA kernels on the B200: forward / backward times per mode and shared-memory budget.

python tools/tune_tile.py [--quick]  -> one JSON line per (shape, loc, dtype, mode, rows).  Design input for the
defaults in csrc/msda_*_tile.cu; the bench values come from bench.py.
"""
from __future__ import annotations

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relation_detr_b200 as rd  # noqa: E402,F401
from relation_detr_b200 import _lib, ops, workloads  # noqa: E402

DEV = "cuda:0"


def time_ms(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    quick = "--quick" in sys.argv
    L_ = _lib.lib()
    cases = [("msda_enc_800x1333_b8", "S", torch.float32), ("msda_enc_800x1333_b8", "U", torch.float32),
             ("msda_enc_800x1333_b8", "S", torch.bfloat16), ("msda_enc_800x1333_b2", "S", torch.float32),
             ("msda_enc_1200x2000_b1", "S", torch.float32)]
    if quick:
        cases = cases[:2]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for name, kind, dtype in cases:
        shape = workloads.MSDA_SHAPES[name]
        inp = workloads.make_msda_inputs(shape, kind, seed=0, device=DEV)
        v = inp["value"].to(dtype)
        go = inp["grad_output"].to(dtype)
        ss, lsi, loc, attn = inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"]
        fb, bb = shape.algorithmic_bytes(2 if dtype == torch.bfloat16 else 4)
        small = shape.batch * shape.S * 256 * v.element_size() < (200 << 20)
        base = None
        for mode, rows in [(1, 0), (2, 256), (2, 384), (2, 512), (2, 640), (2, 800), (2, 1100)]:
            L_.rdetr_msda_set_tile_mode(mode)
            L_.rdetr_msda_set_tile_rows(rows)

            def fwd():
                if small:
                    flush.zero_()
                return ops.msda_forward(v, ss, lsi, loc, attn)

            def bwd():
                if small:
                    flush.zero_()
                return ops.msda_backward(v, ss, lsi, loc, attn, go)

            try:
                tf = time_ms(fwd)
                tb = time_ms(bwd)
            except Exception as e:  # noqa: BLE001
                print(json.dumps(dict(shape=name, loc=kind, dtype=str(dtype), mode=mode, rows=rows, error=str(e)[:200])), flush=True)
                continue
            tflush = time_ms(lambda: flush.zero_()) if small else 0.0
            tf, tb = tf - tflush, tb - tflush
            out = ops.msda_forward(v, ss, lsi, loc, attn)
            gv, gl, ga = ops.msda_backward(v, ss, lsi, loc, attn, go)
            if base is None:
                base = (out.float(), gv.float(), gl, ga)
                diff = None
            else:
                diff = [float((a.float() - b).abs().max() / b.abs().max().clamp(min=1e-30)) for a, b in zip((out, gv, gl, ga), base)]
            print(json.dumps(dict(shape=name, loc=kind, dtype=str(dtype).replace("torch.", ""), mode="flat" if mode == 1 else "tile",
                                  rows=rows, fwd_ms=round(tf, 4), bwd_ms=round(tb, 4), fwd_gbs=round(fb / tf / 1e6, 1),
                                  bwd_gbs=round(bb / tb / 1e6, 1), rel_diff_vs_flat=diff)), flush=True)
    L_.rdetr_msda_set_tile_mode(0)
    L_.rdetr_msda_set_tile_rows(0)


if __name__ == "__main__":
    main()
