"""One-call experiment: parity of the lean MSDA kernels + timings of variants (tuning build).  Output: gpurun_out/r02aa_exp_lean.txt"""
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
        f = t(lambda: ops.msda_forward(*a))
        g = inp["grad_output"].to(dt)
        b = t(lambda: ops.msda_backward(*a, g))
        res.append("%%s/%%s fwd %%.4f bwd %%.4f" %% (kind, "f32" if dt == torch.float32 else "bf16", f, b))
print(os.environ.get("TAG"), " | ".join(res))
''' % ROOT
lib = os.path.join(ROOT, "tools", "librdetr_tune.so")
runs = [
    ("default", {}),
    ("fwd 8ch LDG.256 lean uncapped (v16)", {"RDETR_MSDA_FWD_VARIANT": "16"}),
    ("fwd 8ch LDG.256 lean 64thr/16 (v17)", {"RDETR_MSDA_FWD_VARIANT": "17"}),
    ("fwd 8ch LDG.256 old loop (v18)", {"RDETR_MSDA_FWD_VARIANT": "18"}),
    ("fwd 8ch LDG.256 lean 128thr/8 (v19)", {"RDETR_MSDA_FWD_VARIANT": "19"}),
    ("fwd 8ch LDG.256 lean 32thr (v20)", {"RDETR_MSDA_FWD_VARIANT": "20"}),
    ("1200x2000 default", {"SHAPE": "msda_enc_1200x2000_b1"}),
    ("1200x2000 fwd 8ch LDG.256 (v16)", {"SHAPE": "msda_enc_1200x2000_b1", "RDETR_MSDA_FWD_VARIANT": "16"}),
    ("dec 900 default", {"SHAPE": "msda_dec_900_b8"}),
    ("dec 900 fwd 8ch LDG.256 (v16)", {"SHAPE": "msda_dec_900_b8", "RDETR_MSDA_FWD_VARIANT": "16"}),
    ("enc b2 default", {"SHAPE": "msda_enc_800x1333_b2"}),
    ("enc b2 fwd 8ch LDG.256 (v16)", {"SHAPE": "msda_enc_800x1333_b2", "RDETR_MSDA_FWD_VARIANT": "16"}),
]
for tag, extra in runs:
    env = dict(os.environ, RDETR_OPS_LIB=lib, TAG=tag, **extra)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
