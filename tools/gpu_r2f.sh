#!/bin/bash
# round 2, GPU call F (8 GPUs): config 3 (data-parallel training incl. dp_check), config 4 and config 5 at 8 ranks
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 rc=$?"; }
run 29561 r2f_bench_n8 --steps 30 --warmup 6
python -c "import json;d=json.load(open('gpurun_out/r2f_bench_n8.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks'],d['dp_check'])"; grep -c "teardown clean" gpurun_out/r2f_bench_n8.err
run 29562 r2f_cfg4_n8 --workload config4 --steps 20 --warmup 5 --no-cpu-baseline
python -c "import json;d=json.load(open('gpurun_out/r2f_cfg4_n8.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['dp_check']['ok'])"
run 29563 r2f_cfg5_n8 --workload config5 --steps 4
cut -c1-900 gpurun_out/r2f_cfg5_n8.json
UB_INFER_BANDED=0 run 29564 r2f_cfg5_n8_rr --workload config5 --steps 4
cut -c1-600 gpurun_out/r2f_cfg5_n8_rr.json
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29565 tools/dp_check.py --size 256 --batch 4 > gpurun_out/r2f_dp_check.json 2> gpurun_out/r2f_dp_check.err; echo "dp_check rc=$?"; grep "^{" gpurun_out/r2f_dp_check.json | cut -c1-700; grep -i -m3 "NVLS\|nvls" gpurun_out/r2f_dp_check.err gpurun_out/r2f_dp_check.json | cut -c1-300
