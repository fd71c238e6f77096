"""ctypes + numpy front end of ``librdetr_oracle.so`` (the scalar C restatement).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites the reference lines
it restates in ``rdetr_oracle.c`` / ``rdetr_oracle_impl.inc``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librdetr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc, seconds)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("rdetr_oracle.c", "rdetr_oracle_impl.inc", "lsap_oracle.c"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.rdetr_oracle_set_threads.restype = ctypes.c_int
        _lib.rdetr_oracle_set_threads.argtypes = [ctypes.c_int]
    return _lib


def set_threads(n: int) -> int:
    """n > 0 sets the OpenMP thread count; returns the count in effect."""
    return lib().rdetr_oracle_set_threads(int(n))


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _real(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", np.float32, ctypes.c_float
    if dtype == np.float64:
        return "f64", np.float64, ctypes.c_double
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def msda_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """value [B,S,M,D]; shapes [L,2] (h,w); lsi [L]; loc [B,Nq,M,L,P,2] (x,y); attn [B,Nq,M,L,P]
    -> out [B,Nq,M*D].  Restates ms_deform_im2col_cuda.cuh:226-288."""
    sfx, dt, _ = _real(value.dtype)
    value, loc, attn = _c(value, dt), _c(sampling_locations, dt), _c(attention_weights, dt)
    shapes, lsi = _c(spatial_shapes, np.int64), _c(level_start_index, np.int64)
    B, S, M, D = value.shape
    _, Nq, _, L, P, _ = loc.shape
    out = np.empty((B, Nq, M * D), dtype=dt)
    fn = getattr(lib(), f"rdetr_oracle_msda_forward_{sfx}")
    fn.restype = None
    fn(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), _p(out),
       *(ctypes.c_int(int(x)) for x in (B, S, M, D, L, Nq, P)))
    return out


def msda_backward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, grad_output):
    """-> (grad_value, grad_loc, grad_attn).  Restates ms_deform_im2col_cuda.cuh:76-148, 290-392."""
    sfx, dt, _ = _real(value.dtype)
    value, loc, attn = _c(value, dt), _c(sampling_locations, dt), _c(attention_weights, dt)
    go = _c(grad_output, dt)
    shapes, lsi = _c(spatial_shapes, np.int64), _c(level_start_index, np.int64)
    B, S, M, D = value.shape
    _, Nq, _, L, P, _ = loc.shape
    gv, gl, ga = np.empty_like(value), np.empty_like(loc), np.empty_like(attn)
    fn = getattr(lib(), f"rdetr_oracle_msda_backward_{sfx}")
    fn.restype = None
    fn(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), _p(go), _p(gv), _p(gl), _p(ga),
       *(ctypes.c_int(int(x)) for x in (B, S, M, D, L, Nq, P)))
    return gv, gl, ga


def rel_forward(src_boxes, tgt_boxes, weight, bias, dim_t, scale=100.0, eps=1e-5, attn_mask=None):
    """src [B,N1,4], tgt [B,N2,4] cxcywh; weight [H,8K]; bias [H]; dim_t [K] -> out [B,H,N1,N2].
    Restates relation_transformer.py:481-490,512-532 and position_encoding.py:131-138."""
    sfx, dt, creal = _real(src_boxes.dtype)
    src, tgt = _c(src_boxes, dt), _c(tgt_boxes, dt)
    w, b_, d = _c(weight, dt).reshape(len(bias), -1), _c(bias, dt), _c(dim_t, dt)
    B, N1, _ = src.shape
    N2, H, K = tgt.shape[1], w.shape[0], d.shape[0]
    assert w.shape[1] == 8 * K
    mask = None if attn_mask is None else _c(attn_mask, np.uint8)
    out = np.empty((B, H, N1, N2), dtype=dt)
    fn = getattr(lib(), f"rdetr_oracle_rel_forward_{sfx}")
    fn.restype = None
    fn(_p(src), _p(tgt), _p(w), _p(b_), _p(d), creal(scale), creal(eps),
       _p(mask) if mask is not None else ctypes.c_void_p(0), _p(out),
       *(ctypes.c_int(int(x)) for x in (B, N1, N2, H, K)))
    return out


def rel_backward(src_boxes, tgt_boxes, weight, bias, dim_t, grad_out, scale=100.0, eps=1e-5):
    """-> (grad_weight [H,8K], grad_bias [H]); no gradient reaches the boxes (relation_transformer.py:527)."""
    sfx, dt, creal = _real(src_boxes.dtype)
    src, tgt = _c(src_boxes, dt), _c(tgt_boxes, dt)
    w, b_, d = _c(weight, dt).reshape(len(bias), -1), _c(bias, dt), _c(dim_t, dt)
    go = _c(grad_out, dt)
    B, N1, _ = src.shape
    N2, H, K = tgt.shape[1], w.shape[0], d.shape[0]
    gw, gb = np.empty_like(w), np.empty_like(b_)
    fn = getattr(lib(), f"rdetr_oracle_rel_backward_{sfx}")
    fn.restype = None
    fn(_p(src), _p(tgt), _p(w), _p(b_), _p(d), creal(scale), creal(eps), _p(go), _p(gw), _p(gb),
       *(ctypes.c_int(int(x)) for x in (B, N1, N2, H, K)))
    return gw, gb


def lsap(cost):
    """cost [n_rows, n_cols] -> (row_ind, col_ind) int64, the pairs scipy.optimize.linear_sum_assignment
    returns for the same matrix (hungarian_matcher.py:80).  Raises ValueError like SciPy on an infeasible
    matrix or a NaN / -inf entry.  See lsap_oracle.c for what is restated and how it is pinned."""
    c = np.ascontiguousarray(np.asarray(cost), dtype=np.float64)
    assert c.ndim == 2
    n = min(c.shape)
    rows, cols = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
    fn = lib().rdetr_oracle_lsap
    fn.restype = ctypes.c_int
    rc = fn(_p(c), ctypes.c_int64(c.shape[0]), ctypes.c_int64(c.shape[1]), _p(rows), _p(cols))
    if rc == -1:
        raise ValueError("cost matrix is infeasible")
    if rc == -2:
        raise ValueError("matrix contains invalid numeric entries")
    return rows, cols
