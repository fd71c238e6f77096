"""Times the CTA-size variants of a tuning build of the MSDA kernels at configs[1].

Build: nvcc ... -DRDETR_TUNE_FWD -o tools/librdetr_tune.so relation-detr_b200/csrc/*.cu; variant 0 = shipped default
(forward 64 threads with fp32 capped at 32 registers, backward 128 threads); forward 1/2/3 = 256/128/32 threads,
4 = 64 threads uncapped, 5/6/7 = (64 thr, 24 CTAs), (128, 16), (32, 32); backward 1/2 = 256/64 threads, 3 = 128 threads capped at 40 registers."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, torch
sys.path.insert(0, %r)
from relation_detr_b200 import ops, workloads
shape = workloads.MSDA_SHAPES["msda_enc_800x1333_b8"]
res = []
for kind in ("S", "U"):
    inp = workloads.make_msda_inputs(shape, kind, seed=0, device="cuda:0")
    for dt in (torch.float32, torch.bfloat16):
        v = inp["value"].to(dt)
        a = (v, inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
        for _ in range(3): ops.msda_forward(*a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.msda_forward(*a)
        e1.record(); torch.cuda.synchronize()
        res.append("%%s/%%s %%.4f" %% (kind, "f32" if dt == torch.float32 else "bf16", e0.elapsed_time(e1) / 10))
print("variant", os.environ.get("RDETR_MSDA_FWD_VARIANT"), " ".join(res))
''' % ROOT
code_bwd = r'''
import os, sys, torch
sys.path.insert(0, %r)
from relation_detr_b200 import ops, workloads
shape = workloads.MSDA_SHAPES["msda_enc_800x1333_b8"]
res = []
for kind in ("S", "U"):
    inp = workloads.make_msda_inputs(shape, kind, seed=0, device="cuda:0")
    for dt in (torch.float32, torch.bfloat16):
        v = inp["value"].to(dt)
        a = (v, inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"], inp["grad_output"].to(dt))
        for _ in range(3): ops.msda_backward(*a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.msda_backward(*a)
        e1.record(); torch.cuda.synchronize()
        res.append("%%s/%%s %%.4f" %% (kind, "f32" if dt == torch.float32 else "bf16", e0.elapsed_time(e1) / 10))
print("bwd variant", os.environ.get("RDETR_MSDA_BWD_VARIANT"), " ".join(res))
''' % ROOT
lib = os.path.join(ROOT, "tools", "librdetr_tune.so")
for v in (0, 1, 2, 3, 4, 5, 6, 7):
    env = dict(os.environ, RDETR_MSDA_FWD_VARIANT=str(v), RDETR_OPS_LIB=lib)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
for v in (0, 1, 2, 3):
    env = dict(os.environ, RDETR_MSDA_BWD_VARIANT=str(v), RDETR_OPS_LIB=lib)
    out = subprocess.run([sys.executable, "-c", code_bwd], env=env, capture_output=True, text=True)
    print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
