#!/bin/bash
# round 2, GPU call I: 256-column pair tiles (experiment): parity, sustained, step A/B
mkdir -p gpurun_out
UB_CONV3_PAIR256=1 UB_CASE_TIMEOUT=240 UB_PROBE_OUT=r2i_probe_p256.json timeout 900 python tests/gpu_probe.py conv_fwd_cat_64+64_256,conv_fwd_128_512,conv_dgrad_256_512,conv_fwd_bn_256,layer_enc3b,layer_dec4a,layer_botb > gpurun_out/r2i_probe_p256.log 2>&1; echo "probe rc=$?"
cut -c1-300 gpurun_out/r2i_probe_p256.log
for v in 0 1; do
  UB_CONV3_PAIR256=$v timeout 300 python tools/sustained.py 2 enc3b_fwd enc3b_dgrad enc4b_fwd enc4b_dgrad botb_fwd botb_dgrad > gpurun_out/r2i_sustained_p256_$v.jsonl 2> gpurun_out/r2i_sustained_p256_$v.err; echo "sustained pair256=$v rc=$?"; cat gpurun_out/r2i_sustained_p256_$v.jsonl
done
for v in 0 1 0 1; do
  UB_CONV3_PAIR256=$v timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2i_bench_p256_$v.json 2> gpurun_out/r2i_bench_p256_$v.err
  echo "pair256=$v rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2i_bench_p256_$v.json'));k=d['kernel_ms_per_step'];print(round(d['ms_per_step'],3),round(d['value'],1),d['clocks']['sm_mhz'],d.get('final_loss'),'fwd',k.get('ub_conv3x3_fwd_bn'),'dgrad',k.get('ub_conv3x3_dgrad'),k.get('ub_conv3x3_dgrad_bnred'))")"
done
