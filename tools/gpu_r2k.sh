#!/bin/bash
# round 2, GPU call K (8 GPUs): N = 8 with and without the gradient all-reduce on one box
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29570 + RANDOM % 200)) bench.py --gpus 8 --steps 40 --warmup 6 --no-cpu-baseline > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/$name.json'));c=d['dp_check'];print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks']['sm_mhz'],c.get('ok'),c.get('buckets'),c.get('skipped','')[:20])")"; }
run r2k_n8_noreduce UB_DP_NOREDUCE=1
run r2k_n8_default UB_X=0
timeout 300 python bench.py --no-cpu-baseline --steps 40 --warmup 6 > gpurun_out/r2k_n1.json 2> gpurun_out/r2k_n1.err; echo "n1 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2k_n1.json'));print(round(d['ms_per_step'],3),round(d['value'],1),d['clocks']['sm_mhz'])")"
