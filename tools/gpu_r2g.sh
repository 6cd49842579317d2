#!/bin/bash
# round 2, GPU call G: CTA-pair (cta_group::2) conv3 kernel: parity (kernel cases under UB_CONV3_2CTA=1), sustained rate A/B, step A/B;
# plus the cases added since call E (K > 8 head, label smoothing, graph replay across reallocation)
mkdir -p gpurun_out
UB_CASE_TIMEOUT=300 UB_PROBE_OUT=r2g_probe_new.json timeout 1500 python tests/gpu_probe.py head_k,head_argmax,live_fp32_c3k20,live_bf16_c1k40,live_fp32_c1k2_smooth,graph_two_shapes,algebra_ > gpurun_out/r2g_probe_new.log 2>&1; echo "probe new rc=$?"
cut -c1-330 gpurun_out/r2g_probe_new.log
UB_CONV3_2CTA=1 UB_CASE_TIMEOUT=240 UB_PROBE_OUT=r2g_probe_2cta.json timeout 1500 python tests/gpu_probe.py conv_fwd_128,conv_fwd_cat,conv_dgrad_256,conv_dgrad_split,conv_dgrad_bnred_128_128,conv_dgrad_bnred_256,conv_dgrad_bnred_cat_128,conv_fwd_folded_128,conv_fwd_bn_cat,conv_fwd_bn_256,layer_enc2a,layer_enc3b,layer_dec4a,layer_botb > gpurun_out/r2g_probe_2cta.log 2>&1; echo "probe 2cta rc=$?"
cut -c1-420 gpurun_out/r2g_probe_2cta.log
for v in 0 1; do
  UB_CONV3_2CTA=$v timeout 300 python tools/sustained.py 2 dec2a_fwd dec2a_dgrad enc3b_fwd enc3b_dgrad enc4b_fwd botb_fwd > gpurun_out/r2g_sustained_2cta$v.jsonl 2> gpurun_out/r2g_sustained_2cta$v.err; echo "sustained 2cta=$v rc=$?"; cat gpurun_out/r2g_sustained_2cta$v.jsonl; tail -2 gpurun_out/r2g_sustained_2cta$v.err | cut -c1-300
done
for v in 0 1 0 1; do
  UB_CONV3_2CTA=$v timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2g_bench_2cta$v.json 2> gpurun_out/r2g_bench_2cta$v.err
  echo "2cta=$v rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2g_bench_2cta$v.json'));k=d['kernel_ms_per_step'];print(round(d['ms_per_step'],3),round(d['value'],1),d['clocks']['sm_mhz'],d.get('final_loss'),'fwd',k.get('ub_conv3x3_fwd_bn'),'dgrad',k.get('ub_conv3x3_dgrad'),k.get('ub_conv3x3_dgrad_bnred'))")"
done
