"""Build the reference's own CUDA extension, unmodified, for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY (GPU baseline "the reference kernel recompiled for B200" + a second oracle
with fp64 support).  Sources are compiled where they lie under /root/reference; nothing is copied.
The only addition is a pre-included shim (oracle/ref_cuda_shim.h) restoring a dispatch overload that
current torch dropped.  Takes ~6 minutes (torch headers); the resulting .so is git-ignored but
travels to the GPU box.  Run in the build container:  python -m oracle.build_ref_cuda
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "MultiScaleDeformableAttention"


def build(reference_root: str = "/root/reference", verbose: bool = True):
    from torch.utils.cpp_extension import load

    src = os.path.join(reference_root, "models", "bricks", "ops", "cuda", "ms_deform_attn_cuda.cu")
    if not os.path.exists(src):
        raise RuntimeError(f"reference source not found: {src}")
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    shim = os.path.join(HERE, "ref_cuda_shim.h")
    return load(NAME, sources=[src], extra_cflags=["-O2"], extra_cuda_cflags=["-include", shim, "-O2"],
                build_directory=OUT, verbose=verbose)


def load_prebuilt():
    """Import oracle/_ref/MultiScaleDeformableAttention.so if it exists (GPU box); None otherwise."""
    import importlib.util

    import torch  # noqa: F401  (the extension links libtorch)

    path = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    m = build(*(sys.argv[1:2]))
    print("built", m)
