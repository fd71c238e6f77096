#!/bin/bash
# train-step proxy at 1..N GPUs (N = number of visible GPUs), fp32 / tf32 / bf16
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
: > gpurun_out/train_proxy.jsonl
for prec in "" "--bf16" "--graph --bf16" "--backbone" "--backbone --bf16"; do
  timeout 200 python tools/train_proxy.py $prec 2>/dev/null | grep '^{' >> gpurun_out/train_proxy.jsonl
  for n in 2 4 8; do
    if [ $n -le $N ]; then
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) tools/train_proxy.py $prec 2>gpurun_out/tp_err_$n.log | grep '^{' >> gpurun_out/train_proxy.jsonl
    fi
  done
done
cut -c100-400 gpurun_out/train_proxy.jsonl
