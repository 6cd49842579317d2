#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/gpu_probe.py pool_bwd,head_k,golden,live_bf16,curve > gpurun_out/probe_new.log 2>&1; echo "probe rc=$?"
cut -c1-700 gpurun_out/probe_new.log
for f in 1 0; do UB_FUSE_EW=$f timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/bench_fuse$f.json 2> gpurun_out/bench_fuse$f.err; echo "fuse=$f rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_fuse$f.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks']['sm_mhz'],d['kernel_ms_per_step'].get('ub_bn_bwd_reduce'),d['kernel_ms_per_step'].get('ub_maxpool2x2_bwd_add_bnred'),d['kernel_ms_per_step'].get('ub_head_bwd_apply_bnred'))")"; done
