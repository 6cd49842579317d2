#!/bin/bash
# round 2, call U (2 GPUs): data-parallel bench line of the final build with its dp_check
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_bench_n2.json 2> gpurun_out/r2u_bench_n2.err; echo "bench n2 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2u_bench_n2.json'));print(round(d['ms_per_step'],3),round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['dp_check'])" ; tail -3 gpurun_out/r2u_bench_n2.err
