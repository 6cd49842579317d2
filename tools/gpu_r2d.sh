#!/bin/bash
# round 2, GPU call D (2 GPUs): fused-finalise kernels, data-parallel numerics on live NCCL ranks, N=2 bench with dp_check, banded inference at 2 ranks
mkdir -p gpurun_out
UB_CASE_TIMEOUT=600 UB_PROBE_OUT=r2d_probe.json timeout 1200 python tests/gpu_probe.py fwd_bn,live_bf16_c1k2,golden_c1_k2_bf16,curve_bf16 > gpurun_out/r2d_probe.log 2>&1; echo "probe rc=$?"
cut -c1-600 gpurun_out/r2d_probe.log
timeout 600 python -m pytest tests/test_dp_gpu.py -q -m gpu > gpurun_out/r2d_dp_pytest.log 2>&1; echo "dp pytest rc=$?"; tail -5 gpurun_out/r2d_dp_pytest.log | cut -c1-1500
for cfg in "1 0" "0 0" "1 1"; do
  set -- $cfg
  UB_FUSE_FINALIZE=$1 UB_FUSE_EW=$2 timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2d_bench_fin$1_ew$2.json 2> gpurun_out/r2d_bench_fin$1_ew$2.err
  echo "fin=$1 ew=$2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2d_bench_fin$1_ew$2.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks']['sm_mhz'],d['gpu_launches'],d.get('final_loss'))")"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 30 --warmup 6 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
echo "n2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2d_bench_n2.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['dp_check'])")"; grep -i "teardown" gpurun_out/r2d_bench_n2.err | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --workload config5 --steps 3 > gpurun_out/r2d_cfg5_n2.json 2> gpurun_out/r2d_cfg5_n2.err
echo "cfg5 n2 rc=$? $(cut -c1-700 gpurun_out/r2d_cfg5_n2.json)"; tail -3 gpurun_out/r2d_cfg5_n2.err
UB_INFER_BANDED=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --workload config5 --steps 3 > gpurun_out/r2d_cfg5_n2_rr.json 2> gpurun_out/r2d_cfg5_n2_rr.err
echo "cfg5 n2 round-robin rc=$? $(cut -c1-500 gpurun_out/r2d_cfg5_n2_rr.json)"
