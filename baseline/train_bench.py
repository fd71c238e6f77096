"""Relation-DETR R50 training-step throughput (BASELINE.json configs[3]: 800x1333, batch 2 per GPU, DDP).

The model is the reference's own ``RelationDETR`` built from its own classes (``baseline/refmodel.py``); the step
is the reference's (``baseline/train_step.py``).  Three operator paths:

  ours            ``relation_detr_b200.install.install()`` before the model is built: B200 MSDA / relation / matcher
  ours_fused_attention  the same plus ``install(fused_attention=True)``: the decoder's self-attention generates the relation
                  bias inside the attention kernel
  ours_all        everything: also ``install(fused_memory=True)`` (memory_fusion's input Linear as the K-split tcgen05 GEMM, row N4;
                  active under autocast / allow_tf32) and ``install(fused_topk=True)`` (the two-stage selection's torch.topk as
                  rdetr_topk_rows)
  reference       the unmodified reference as it runs on this image (its extension does not build -> grid_sample path,
                  eager relation embedding, SciPy matcher with one device->host copy per prediction set)
  reference_cuda  the unmodified reference with its own CUDA kernel (``oracle/_ref``, sources untouched, sm_100a)

``run(...)`` is called by ``bench.py`` on every rank (DDP over NCCL when world > 1); timings are CUDA events on the
current stream, max over ranks; images / targets are synthetic and fixed (seed = rank).
"""
from __future__ import annotations

import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _max_over_ranks(x: float, dev) -> float:
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x


def run(path: str, precision: str, steps: int, warmup: int, batch_per_gpu: int = 2, height: int = 800, width: int = 1333,
        boxes_per_image: int = 10, enc_layers: int = 6, dec_layers: int = 6, log_sync: bool = False, profile_share: bool = False,
        model_name: str = "r50") -> dict:
    """``model_name="focal_l"`` = BASELINE configs[4] (FocalNet-L, 5 levels; call it with height=1200, width=2000, batch 1)."""
    import torch
    import torch.distributed as dist

    from baseline import refmodel, train_step
    from relation_detr_b200 import install as rinstall

    dev = torch.device("cuda", torch.cuda.current_device())
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    refmodel.activate()
    rinstall.uninstall()
    graphed = path == "ours_graphed"   # "ours" + CUDA-graph capture of backbone / encoder / decoder passes (relation_detr_b200.graphs)
    if graphed:
        path = "ours"
    if path in ("ours", "ours_fused_attention", "ours_all"):
        report = rinstall.install(fused_attention=(path in ("ours_fused_attention", "ours_all")), fused_memory=(path == "ours_all"),
                                  fused_topk=(path == "ours_all"))
        assert not report.skipped
        ext = "rdetr"
    else:
        ext = refmodel.set_reference_extension("prebuilt" if path == "reference_cuda" else "none")
        if path == "reference_cuda" and ext != "prebuilt":
            return {"unavailable": "oracle/_ref (the reference's own CUDA kernel, prebuilt) is not present"}
    torch.manual_seed(0)
    if model_name == "focal_l":
        model, _ = refmodel.build_relation_detr_focal_l(enc_layers=enc_layers, dec_layers=dec_layers)
    else:
        model, _ = refmodel.build_relation_detr_r50(enc_layers=enc_layers, dec_layers=dec_layers)
    if path not in ("ours_fused_attention", "ours_all"):  # the lazy relation-bias hand-over is a class-level switch: it stays on while the model runs
        rinstall.uninstall()
    model = model.to(dev).train()
    n_params = sum(p.numel() for p in model.parameters())
    train_model = model
    amp = torch.bfloat16 if precision == "bf16" else None
    images, targets = refmodel.synthetic_batch(batch_per_gpu, dev, seed=rank, height=height, width=width, boxes_per_image=boxes_per_image)
    handle = None
    if graphed:
        from relation_detr_b200 import graphs

        handle = graphs.capture_static_parts(model, images, targets, autocast_dtype=amp)
    if world > 1:
        train_model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], find_unused_parameters=False)
    opt = train_step.build_optimizer(model)

    def step():
        return train_step.train_step(train_model, images, targets, opt, autocast_dtype=amp, log_sync=log_sync, autocast_cache=not graphed)

    for _ in range(warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    ms = _max_over_ranks(e0.elapsed_time(e1) / steps, dev)
    res = {"path": "ours_graphed" if graphed else path, "model": model_name, "precision": precision, "msda_path": ext, "ms_per_step": round(ms, 2), "host_ms_per_step": round(wall_ms, 2),
           "imgs_per_s": round(world * batch_per_gpu / ms * 1e3, 2), "global_batch": world * batch_per_gpu, "steps": steps, "warmup": warmup,
           "loss": round(float(loss), 4), "params_M": round(n_params / 1e6, 2),
           "peak_mem_GB": round(torch.cuda.max_memory_allocated(dev) / 2**30, 2)}
    if profile_share and rank == 0:
        try:
            from torch.profiler import ProfilerActivity, profile

            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step()
                torch.cuda.synchronize()
            tot = sum(e.device_time_total for e in prof.key_averages())
            ours = sum(e.device_time_total for e in prof.key_averages() if "rdetr::" in e.key)
            nccl = sum(e.device_time_total for e in prof.key_averages() if "nccl" in e.key.lower())
            res["kernel_time_share"] = {"rdetr_kernels": round(ours / max(tot, 1), 4), "nccl": round(nccl / max(tot, 1), 4),
                                        "device_busy_ms": round(tot / 1e3, 2)}
        except Exception as e:  # noqa: BLE001  (diagnostic only)
            res["kernel_time_share"] = {"error": f"{type(e).__name__}: {e}"[:160]}
    if handle is not None:
        res["graphed_parts"] = list(handle.parts)
        handle.release()
    del opt, train_model, model
    rinstall.uninstall()
    torch.cuda.empty_cache()
    return res


def gpu_inference(path: str = "ours", precision: str = "fp32", passes: int = 10, warmup: int = 3, height: int = 800, width: int = 1333) -> dict:
    """The recipe of tools/benchmark_model.py:26-61 (eval, ``eval_transform = None``, batch 1, inference_mode) on the GPU:
    latency of one whole-model forward for the given operator path (``ours`` / ``reference`` / ``reference_cuda``)."""
    import torch

    from baseline import refmodel
    from relation_detr_b200 import install as rinstall

    dev = torch.device("cuda", torch.cuda.current_device())
    refmodel.activate()
    rinstall.uninstall()
    graphed = path == "ours_graphed"
    if path in ("ours", "ours_graphed"):
        rinstall.install()
        ext = "rdetr"
    else:
        ext = refmodel.set_reference_extension("prebuilt" if path == "reference_cuda" else "none")
        if path == "reference_cuda" and ext != "prebuilt":
            return {"unavailable": "oracle/_ref (the reference's own CUDA kernel, prebuilt) is not present"}
    torch.manual_seed(0)
    model, _ = refmodel.build_relation_detr_r50()
    rinstall.uninstall()
    model.eval_transform = None
    model = model.to(dev).eval()
    image = torch.randn(3, height, width, device=dev)
    amp = torch.bfloat16 if precision == "bf16" else None
    handle = None
    if graphed:
        from relation_detr_b200 import graphs

        handle = graphs.capture_static_parts(model, (image,), None, autocast_dtype=amp)
    with torch.inference_mode(), torch.autocast("cuda", dtype=amp, enabled=amp is not None, cache_enabled=not graphed):
        for _ in range(warmup):
            model((image,))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(passes):
            model((image,))
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / passes
    ms = e0.elapsed_time(e1) / passes
    if handle is not None:
        handle.release()
    res = {"path": path, "precision": precision, "msda_path": ext, "ms_per_image": round(ms, 2), "host_ms_per_image": round(wall, 2),
           "imgs_per_s": round(1e3 / ms, 2), "passes": passes, "peak_mem_GB": round(torch.cuda.max_memory_allocated(dev) / 2**30, 2)}
    del model
    torch.cuda.empty_cache()
    return res


def cpu_inference(passes: int = 3, height: int = 800, width: int = 1333) -> dict:
    """BASELINE.json configs[0]: whole Relation-DETR R50, eval, batch 1, synthetic 800x1333 image, on the host cores
    through the reference's own CPU path (recipe of tools/benchmark_model.py:26-61: eval_transform = None,
    inference_mode, 1 warm-up + `passes` timed)."""
    import torch

    from baseline import refmodel
    from relation_detr_b200 import install as rinstall

    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    refmodel.activate()
    rinstall.uninstall()
    refmodel.set_reference_extension("none")
    torch.manual_seed(0)
    model, _ = refmodel.build_relation_detr_r50()
    model.eval_transform = None
    model.eval()
    image = torch.randn(3, height, width)
    with torch.inference_mode():
        model((image,))
        ts = []
        for _ in range(passes):
            t0 = time.perf_counter()
            model((image,))
            ts.append(time.perf_counter() - t0)
    return {"workload": "Relation-DETR R50 800x1333 inference, batch 1, random init, synthetic image (BASELINE configs[0])",
            "s_per_image": round(min(ts), 3), "mean_s_per_image": round(sum(ts) / len(ts), 3), "imgs_per_s": round(1.0 / min(ts), 4),
            "passes": passes, "cores": torch.get_num_threads(), "kind": "reference"}


if __name__ == "__main__":
    import json

    import torch

    path = sys.argv[1] if len(sys.argv) > 1 else "ours"
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    if path == "cpu":
        print(json.dumps(cpu_inference(1)))
    elif len(sys.argv) > 3 and sys.argv[3] == "infer":
        torch.cuda.set_device(0)
        print(json.dumps(gpu_inference(path, prec)))
    elif len(sys.argv) > 3 and sys.argv[3] == "focal_l":   # BASELINE configs[4]
        torch.cuda.set_device(0)
        print(json.dumps(run(path, prec, 3, 2, batch_per_gpu=1, height=1200, width=2000, profile_share=True, model_name="focal_l")))
    else:
        torch.cuda.set_device(0)
        print(json.dumps(run(path, prec, 5, 2, profile_share=True)))
