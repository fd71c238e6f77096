"""Record the DRAM traffic of one MSDA backward call (kernel + grad_value zero-fill) from an ncu capture into
profiles/traffic.json, which bench.py quotes as ``roofline.traffic`` (never a constant in code).

    python tools/record_traffic.py gpurun_out/prof_X.ncu-rep [commit]
The capture must be `ncu --set full -k regex:msda_bwd_kernel` of `tools/profile_ops.py msda` at configs[1], loc S.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(rep, commit=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        v = float(r[idx[key]].replace(",", ""))
        u = units[idx[key]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    per = {}
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        per.setdefault(name, []).append(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
    bwd = [k for k in per if "msda_bwd_kernel" in k]
    if not bwd:
        raise SystemExit(f"no msda_bwd_kernel launch in {rep}: {list(per)}")
    kernel = sum(per[bwd[0]]) / len(per[bwd[0]])
    B, S, M, D = 8, 22323, 8, 32
    memset = B * S * M * D * 4  # cudaMemsetAsync of grad_value inside rdetr_msda_backward: one pass of writes
    commit = commit or subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    out = {"dram_bytes_per_call": int(kernel + memset), "kernel_dram_bytes": int(kernel), "memset_bytes": memset,
           "kernel": bwd[0], "launches_averaged": len(per[bwd[0]]), "commit": commit, "capture": os.path.basename(rep),
           "how": "dram__bytes_read.sum + dram__bytes_write.sum of the backward kernel (ncu --set full, configs[1] loc S) + the "
                  "grad_value zero-fill the same call enqueues (write-only, its size)"}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(*sys.argv[1:3])
