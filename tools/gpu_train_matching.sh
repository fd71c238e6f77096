#!/bin/bash
# train-step proxy on one GPU with per-layer bipartite matching: none / the reference's way / this repository's
mkdir -p gpurun_out
: > gpurun_out/train_matching.jsonl
for prec in "" "--bf16"; do
  for how in none scipy device; do
    timeout 200 python tools/train_proxy.py $prec --matching $how 2>gpurun_out/tm_err.log | grep '^{' >> gpurun_out/train_matching.jsonl || tail -5 gpurun_out/tm_err.log
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/train_matching.jsonl"):
    d = json.loads(l); print(d["precision"], d["matching"], d["ms_per_step"], d["imgs_per_s"], d["loss"])
PY
