#!/bin/bash
# First GPU call of the next round: everything that was written after the round-1 GPU budget ran out.
#   1. pending parity cases (BatchNorm fold kernels, y-less pool, folded head, whole-graph fold, banded inference, known answers)
#   2. A/B of the training step with and without the fold (UB_FOLD_BN)
# usage (1 GPU): tools/gpu_pending.sh          8 GPUs, banded inference: see the torchrun lines at the end
mkdir -p gpurun_out
timeout 900 python tests/gpu_probe.py --pending > gpurun_out/probe_pending.log 2>&1; echo "pending rc=$?"
cut -c1-400 gpurun_out/probe_pending.log
for f in 0 1; do
  UB_FOLD_BN=$f timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/bench_fold$f.json 2> gpurun_out/bench_fold$f.err
  echo "fold=$f rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_fold$f.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks']['sm_mhz'],d['final_loss'])")"
done
# multi-GPU (run separately with gpurun --gpus 8):
#   torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/bench_infer.py --size 20000
#   UB_INFER_BANDED=1 torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/bench_infer.py --size 20000
