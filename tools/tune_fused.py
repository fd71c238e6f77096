"""Times the fused-prologue forward under the CTA/register variants of a tuning build."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, torch
sys.path.insert(0, %r)
import bench
from relation_detr_b200 import ops, workloads
shape = workloads.MSDA_SHAPES["msda_enc_800x1333_b8"]
r = []
for dt in (torch.float32, torch.bfloat16):
    for m in (False, True):
        t = bench.time_msda_fused(torch, ops, workloads, shape, 10, 3, dt, m)
        r.append("%%s mask=%%d fwd %%.4f bwd %%.4f" %% ("f32" if dt == torch.float32 else "bf16", m, t["fwd_ms"], t["bwd_ms"]))
print("variant", os.environ.get("RDETR_MSDA_FWD_VARIANT"), " | ".join(r))
''' % ROOT
lib = os.path.join(ROOT, "tools", "librdetr_tune.so")
for v in (0, 4, 5, 2):
    env = dict(os.environ, RDETR_MSDA_FWD_VARIANT=str(v), RDETR_OPS_LIB=lib)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:])
