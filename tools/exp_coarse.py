"""Backward timings with the coarse-level kernel off (RDETR_MSDA_COARSE=1) and on (=2).  Output: one line per run."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, torch
sys.path.insert(0, %r)
from relation_detr_b200 import ops, workloads
shape = workloads.MSDA_SHAPES[os.environ.get("SHAPE", "msda_enc_800x1333_b8")]
res = []
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for kind in ("S", "U"):
    inp = workloads.make_msda_inputs(shape, kind, seed=0, device="cuda:0")
    for dt in (torch.float32, torch.bfloat16):
        v = inp["value"].to(dt)
        a = (v, inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
        g = inp["grad_output"].to(dt)
        b = t(lambda: ops.msda_backward(*a, g))
        res.append("%%s/%%s bwd %%.4f" %% (kind, "f32" if dt == torch.float32 else "bf16", b))
print(os.environ.get("TAG"), " | ".join(res))
''' % ROOT
runs = []
for shape in ("msda_enc_800x1333_b8", "msda_enc_800x1333_b2", "msda_enc_800x1333_b1", "msda_enc_1200x2000_b1", "msda_dec_900_b8"):
    for mode in ("1", "2"):
        runs.append(("%s coarse=%s" % (shape, "off" if mode == "1" else "on"), {"SHAPE": shape, "RDETR_MSDA_COARSE": mode}))
for tag, extra in runs:
    env = dict(os.environ, TAG=tag, **extra)
    try:
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
    except subprocess.TimeoutExpired:
        print(tag, "TIMEOUT", flush=True)
