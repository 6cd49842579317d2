#!/bin/bash
mkdir -p gpurun_out
UB_CASE_TIMEOUT=300 UB_PROBE_OUT=r2m_probe.json timeout 900 python tests/gpu_probe.py conv_first_c,conv_first_tiles,augment_c,live_fp32_c6k3,live_fp32_c1k2_smooth,live_fp32_c3k8 > gpurun_out/r2m_probe.log 2>&1; echo "probe rc=$?"
cut -c1-260 gpurun_out/r2m_probe.log
for v in 1 0 1 0; do
  UB_FUSE_FINALIZE=$v timeout 300 python bench.py --no-cpu-baseline --steps 40 --warmup 6 > gpurun_out/r2m_bench_fin$v.json 2> gpurun_out/r2m_bench_fin$v.err
  echo "fin=$v rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2m_bench_fin$v.json'));k=d['kernel_ms_per_step'];print(round(d['ms_per_step'],3),round(d['value'],1),d['clocks']['sm_mhz'],'deconv_fwd',k.get('ub_deconv2x2_fwd_bn'),k.get('ub_deconv2x2_fwd'),'finalize',k.get('ub_bn_finalize'))")"
done
timeout 300 compute-sanitizer --tool memcheck python tests/gpu_probe.py --one adam > gpurun_out/r2m_sanitize_adam.log 2>&1; echo "sanitize rc=$?"; tail -12 gpurun_out/r2m_sanitize_adam.log | cut -c1-300
