"""ctypes binding of ``librdetr_ops.so`` (the C ABI declared in ``include/rdetr_ops.h``).

There is no CPU path and no fallback: if the library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RDETR_OPS_LIB", os.path.join(PKG_DIR, "librdetr_ops.so"))

ABI_VERSION = 1  # RDETR_ABI_VERSION in include/rdetr_ops.h
DTYPE_F32, DTYPE_BF16 = 0, 1
REL_EXACT, REL_FAST = 0, 1

# every symbol include/rdetr_ops.h declares: name -> (restype, argtypes)
_SIGNATURES = {
    "rdetr_abi_version": (c_int, []),
    "rdetr_last_error": (c_char_p, []),
    "rdetr_msda_forward": (c_int, [c_void_p] * 6 + [c_int] * 8 + [c_void_p]),
    "rdetr_msda_backward_workspace_bytes": (c_size_t, [c_int] * 8),
    "rdetr_msda_backward": (c_int, [c_void_p] * 9 + [c_int] * 8 + [c_void_p, c_size_t, c_void_p]),
    "rdetr_msda_set_tile_mode": (c_int, [c_int]),
    "rdetr_msda_set_coarse_mode": (c_int, [c_int]),
    "rdetr_msda_set_bf16_scatter": (c_int, [c_int]),
    "rdetr_msda_set_tile_rows": (c_int, [c_int]),
    "rdetr_msda_fused_forward": (c_int, [c_void_p] * 8 + [c_int] * 9 + [c_void_p]),
    "rdetr_msda_fused_backward": (c_int, [c_void_p] * 11 + [c_int] * 9 + [c_void_p, c_size_t, c_void_p]),
    "rdetr_relation_workspace_bytes": (c_size_t, [c_int] * 4),
    "rdetr_relation_forward": (c_int, [c_void_p] * 5 + [c_float, c_float] + [c_void_p] * 3 + [c_int] * 5 + [c_void_p, c_size_t, c_void_p]),
    "rdetr_relation_backward": (c_int, [c_void_p] * 3 + [c_float, c_float] + [c_void_p] * 4 + [c_int] * 5 + [c_void_p, c_size_t, c_void_p]),
    "rdetr_relation_attention_workspace_bytes": (c_size_t, [c_int] * 4),
    "rdetr_relation_attention_forward": (c_int, [c_void_p] * 8 + [c_float, c_float] + [c_void_p] * 3 + [c_int] * 4 + [c_void_p, c_size_t, c_void_p]),
    "rdetr_relation_attention_backward": (c_int, [c_void_p] * 8 + [c_float, c_float] + [c_void_p] * 9 + [c_int] * 4 + [c_void_p, c_size_t, c_void_p]),
    "rdetr_memory_fusion_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int, c_void_p]),
    "rdetr_two_stage_workspace_bytes": (c_size_t, [c_int, c_int]),
    "rdetr_topk_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rdetr_two_stage_select": (c_int, [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p] * 4 + [c_size_t, c_void_p]),
    "rdetr_two_stage_select_backward": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p] * 3),
    "rdetr_lsap_workspace_bytes": (c_size_t, [c_void_p, c_void_p, c_int]),
    "rdetr_lsap_solve": (c_int, [c_void_p] * 6 + [c_int, c_void_p, c_size_t, c_void_p]),
    "rdetr_match_cost": (c_int, [c_void_p] * 7 + [c_int, c_float, c_float, c_float, c_double, c_double, c_int, c_void_p]),
    "rdetr_diag_gather_rows": (c_int, [c_void_p, c_longlong, c_int, c_void_p, ctypes.POINTER(c_longlong), c_void_p]),
    "rdetr_diag_red_rows": (c_int, [c_void_p, c_longlong, c_int, ctypes.POINTER(c_longlong), c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()


class RdetrOpsError(RuntimeError):
    """A C-ABI call returned non-zero (the reference only printf's launch errors, cuh:937-941)."""


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build it with `python -m relation_detr_b200.build` "
                        "(nvcc, sm_100a). relation-detr_b200 has no CPU or PyTorch fallback."
                    )
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)  # AttributeError if the .so does not export it
                    fn.restype = res
                    fn.argtypes = args
                if handle.rdetr_abi_version() != ABI_VERSION:
                    raise RuntimeError(f"{LIB_PATH} reports ABI version {handle.rdetr_abi_version()}, this package binds "
                                       f"version {ABI_VERSION}: rebuild with `python -m relation_detr_b200.build --force`")
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().rdetr_last_error()
        raise RdetrOpsError(f"{what} failed (code {rc}): {msg.decode(errors='replace') if msg else ''}")
