#!/bin/bash
# bench.py + train proxy at the full GPU count of the box
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 --quick > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_n$N.json; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "cpu", d["cpu_baseline"])
PY
for prec in "" "--bf16" "--backbone" "--backbone --bf16"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 tools/train_proxy.py $prec 2>gpurun_out/tp_err_$N.log | grep '^{' | tee -a gpurun_out/train_proxy_n$N.jsonl | cut -c100-400
done
