"""Why the coarse-level kernel loses: each kernel alone, both, the scatter kernel held back, per level class.
Needs the tuning build (the RDETR_COARSE_* knobs are compiled out of the shipped library):
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O2 -shared -DRDETR_TUNE_FWD \
         -I include -I relation-detr_b200/csrc -o tools/librdetr_tune.so relation-detr_b200/csrc/*.cu
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tools", "librdetr_tune.so")
code = r'''
import os, sys, torch
sys.path.insert(0, %r)
from relation_detr_b200 import ops, workloads
shape = workloads.MSDA_SHAPES[os.environ.get("SHAPE", "msda_enc_800x1333_b8")]
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
inp = workloads.make_msda_inputs(shape, "S", seed=0, device="cuda:0")
a = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"], inp["grad_output"])
print(os.environ.get("TAG"), "S/f32 bwd %%.4f" %% t(lambda: ops.msda_backward(*a)))
''' % ROOT
runs = [
    ("on, K1 only (no K2), default carveout", {"RDETR_MSDA_COARSE": "2", "RDETR_COARSE_SKIP": "3", "RDETR_COARSE_CARVEOUT": "0"}),
    ("on, default carveout", {"RDETR_MSDA_COARSE": "2", "RDETR_COARSE_CARVEOUT": "0"}),
    ("off", {"RDETR_MSDA_COARSE": "1"}),
    ("on", {"RDETR_MSDA_COARSE": "2"}),
    ("on, K2 only (both classes)", {"RDETR_MSDA_COARSE": "2", "RDETR_COARSE_ONLY": "1"}),
    ("on, K2 only, big class only", {"RDETR_MSDA_COARSE": "2", "RDETR_COARSE_ONLY": "1", "RDETR_COARSE_SKIP": "2"}),
    ("on, K2 only, small class only", {"RDETR_MSDA_COARSE": "2", "RDETR_COARSE_ONLY": "1", "RDETR_COARSE_SKIP": "1"}),
    ("on, K1 only (no K2)", {"RDETR_MSDA_COARSE": "2", "RDETR_COARSE_SKIP": "3"}),
    ("on, K1 delayed 20us", {"RDETR_MSDA_COARSE": "2", "RDETR_COARSE_DELAY_US": "20"}),
    ("on, K1 delayed 50us", {"RDETR_MSDA_COARSE": "2", "RDETR_COARSE_DELAY_US": "50"}),
    ("on, K1 delayed 50us, big only", {"RDETR_MSDA_COARSE": "2", "RDETR_COARSE_DELAY_US": "50", "RDETR_COARSE_SKIP": "2"}),
    ("on, K1 delayed 50us, small only", {"RDETR_MSDA_COARSE": "2", "RDETR_COARSE_DELAY_US": "50", "RDETR_COARSE_SKIP": "1"}),
]
for tag, extra in runs:
    env = dict(os.environ, TAG=tag, RDETR_OPS_LIB=LIB, **extra)
    try:
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
    except subprocess.TimeoutExpired:
        print(tag, "TIMEOUT", flush=True)
