#!/bin/bash
# round 2, call N: row-streaming conv3 kernel for the 64-output-channel layers (UB_CONV3_ROWS): parity, sustained rate, step A/B
mkdir -p gpurun_out
UB_CONV3_ROWS=2 UB_CASE_TIMEOUT=120 UB_PROBE_OUT=r2n_probe.json timeout 600 python tests/gpu_probe.py rows_,conv_first_tiles_c5,layer_enc1b,layer_dec1a,layer_enc2a,conv_fwd_bn_64_64_big > gpurun_out/r2n_probe.log 2>&1; echo "probe rc=$?"
cut -c1-400 gpurun_out/r2n_probe.log
for v in 0 2; do
  UB_CONV3_ROWS=$v timeout 120 python tools/sustained.py 1.5 enc1b_fwd enc1b_dgrad dec1a_fwd enc2a_dgrad > gpurun_out/r2n_sustained_rows$v.jsonl 2> gpurun_out/r2n_sustained_rows$v.err; echo "sustained rows=$v rc=$?"; cat gpurun_out/r2n_sustained_rows$v.jsonl; tail -3 gpurun_out/r2n_sustained_rows$v.err
done
for v in 0 1 2 0 2; do
  UB_CONV3_ROWS=$v timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2n_bench_rows$v.json 2> gpurun_out/r2n_bench_rows$v.err
  echo "rows=$v rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2n_bench_rows$v.json'));k=d['kernel_ms_per_step'];print(round(d['ms_per_step'],3),round(d['value'],1),d['clocks']['sm_mhz'],'fwd',k.get('ub_conv3x3_fwd_bn'),'dgrad',k.get('ub_conv3x3_dgrad'),'loss',d['final_loss'],'roof',round(d['roofline']['frac'],3))")"
done
