"""Host-side profile of one real-model training step (ours, bf16 or fp32): top CPU ops, launch and sync counts.
python tools/profile_train_host.py [fp32|bf16] [ours|reference|ours_graphed]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import refmodel, train_step  # noqa: E402
from relation_detr_b200 import install as rinstall  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
path = sys.argv[2] if len(sys.argv) > 2 else "ours"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
refmodel.activate()
graphed = path == "ours_graphed"
if path in ("ours", "ours_graphed"):
    rinstall.install()
torch.manual_seed(0)
model, _ = refmodel.build_relation_detr_r50()
rinstall.uninstall()
model = model.to(dev).train()
opt = train_step.build_optimizer(model)
images, targets = refmodel.synthetic_batch(2, dev, seed=0)
amp = torch.bfloat16 if prec == "bf16" else None
if graphed:
    from relation_detr_b200 import graphs

    handle = graphs.capture_static_parts(model, images, targets, autocast_dtype=amp)
    print("graphed parts:", handle.parts)
for _ in range(3):
    train_step.train_step(model, images, targets, opt, autocast_dtype=amp, autocast_cache=not graphed)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile, record_function  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=False) as prof:
    with record_function("STEP"):
        ctx = torch.autocast("cuda", dtype=amp, cache_enabled=not graphed) if amp else torch.autocast("cuda", enabled=False)
        with ctx:
            with record_function("FORWARD+LOSS"):
                ld = model(images, targets)
                loss = sum(ld.values())
        opt.zero_grad()
        with record_function("BACKWARD"):
            loss.backward()
        with record_function("CLIP"):
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
        with record_function("OPT"):
            opt.step()
    torch.cuda.synchronize()
ka = prof.key_averages()
print(ka.table(sort_by="self_cpu_time_total", row_limit=28, max_name_column_width=60))
launches = sum(e.count for e in ka if e.key in ("cudaLaunchKernel", "cuLaunchKernel", "cudaLaunchKernelExC", "cuLaunchKernelEx", "cudaGraphLaunch"))
syncs = {e.key: e.count for e in ka if "Synchronize" in e.key or e.key in ("aten::item", "aten::_local_scalar_dense", "cudaMemcpyAsync")}
print("kernel launches:", launches, "sync-ish:", syncs)
for name in ("STEP", "FORWARD+LOSS", "BACKWARD", "CLIP", "OPT"):
    for e in ka:
        if e.key == name:
            print(f"{name}: cpu {e.cpu_time_total / 1e3:.1f} ms, device {e.device_time_total / 1e3:.1f} ms")
