#!/bin/bash
# round 2, call Q: StepPipeline (e2e), deconv dgrad with fused reduction (UB_FUSE_RED_DECONV), dec1a dgrad with fused reduction (UB_FUSE_RED64=2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "deconv_ or rows_dgrad_bnred or conv_dgrad_bnred or layer_enc1b or layer_dec1a" > gpurun_out/r2q_pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2q_pytest_kernels.log
UB_FUSE_RED64=2 timeout 200 python -m pytest tests/test_kernels_gpu.py -q -x -k "layer_dec1a" > gpurun_out/r2q_pytest_dec1a_red.log 2>&1; echo "pytest dec1a red rc=$?"; tail -3 gpurun_out/r2q_pytest_dec1a_red.log
UB_CASE_TIMEOUT=200 UB_PROBE_OUT=r2q_probe.json timeout 600 python tests/gpu_probe.py step_pipeline,reddeconv_wellcond,nored64_wellcond > gpurun_out/r2q_probe.log 2>&1; echo "probe rc=$?"; cut -c1-300 gpurun_out/r2q_probe.log
for cfg in "1 0" "2 1" "2 0" "1 0" "2 1"; do
  set -- $cfg
  UB_FUSE_RED64=$1 UB_FUSE_RED_DECONV=$2 timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2q_bench_$1_$2.json 2> gpurun_out/r2q_bench_$1_$2.err
  echo "red64=$1 deconv=$2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2q_bench_$1_$2.json'));k=d['kernel_ms_per_step'];e=d['e2e'];print(round(d['ms_per_step'],3),round(d['value'],1),'e2e',round(e['ms_per_step'],3),round(e['value'],1),'serial',round(e['serial_ms_per_step'],3),d['clocks']['sm_mhz'],'bnred',k.get('ub_conv3x3_dgrad_bnred'),'reduce',k.get('ub_bn_bwd_reduce'),'dec_dgrad',k.get('ub_deconv2x2_dgrad'),k.get('ub_deconv2x2_dgrad_bnred'),'loss',d['final_loss'])")"; tail -2 gpurun_out/r2q_bench_$1_$2.err
done
