#!/bin/bash
mkdir -p gpurun_out
UB_VERBOSE=1 UB_CASE_TIMEOUT=900 UB_PROBE_OUT=r2c_probe.json timeout 1500 python tests/gpu_probe.py train_driver,tiled_inference_bf16,config1_refdata,nofold_,wgrad_folded,wellcond_bf16 > gpurun_out/r2c_probe.log 2>&1; echo "probe rc=$?"
cut -c1-900 gpurun_out/r2c_probe.log
timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2c_bench.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks'],d['roofline']['frac'])")"
timeout 600 python tools/sustained.py 3 > gpurun_out/r2c_sustained.jsonl 2> gpurun_out/r2c_sustained.err; echo "sustained rc=$?"; cat gpurun_out/r2c_sustained.jsonl
