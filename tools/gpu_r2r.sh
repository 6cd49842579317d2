#!/bin/bash
# round 2, call R: dec1a dgrad as two row-streaming launches (UB_CONV3_ROWS=3) with the fused reduction of up1 (UB_FUSE_RED64=2); deconv dgrad bnred fix
mkdir -p gpurun_out
UB_CONV3_ROWS=3 UB_FUSE_RED64=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "deconv_dgrad_bnred or rows_dgrad or conv_dgrad or layer_dec1a" > gpurun_out/r2r_pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2r_pytest_kernels.log | cut -c1-300
UB_CONV3_ROWS=3 UB_CASE_TIMEOUT=200 UB_PROBE_OUT=r2r_probe.json timeout 600 python tests/gpu_probe.py reddeconv_wellcond > gpurun_out/r2r_probe.log 2>&1; echo "probe rc=$?"; cut -c1-300 gpurun_out/r2r_probe.log
UB_CONV3_ROWS=3 timeout 60 python tools/sustained.py 1.5 dec1a_dgrad > gpurun_out/r2r_sustained.jsonl 2> gpurun_out/r2r_sustained.err; cat gpurun_out/r2r_sustained.jsonl
for cfg in "2 1" "3 2" "3 1" "2 1" "3 2"; do
  set -- $cfg
  UB_CONV3_ROWS=$1 UB_FUSE_RED64=$2 timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2r_bench_$1_$2.json 2> gpurun_out/r2r_bench_$1_$2.err
  echo "rows=$1 red64=$2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2r_bench_$1_$2.json'));k=d['kernel_ms_per_step'];e=d['e2e'];print(round(d['ms_per_step'],3),round(d['value'],1),'e2e',round(e['ms_per_step'],3),round(e['value'],1),d['clocks']['sm_mhz'],'dgrad',k.get('ub_conv3x3_dgrad'),'bnred',k.get('ub_conv3x3_dgrad_bnred'),'reduce',k.get('ub_bn_bwd_reduce'),'loss',d['final_loss'])")"; tail -2 gpurun_out/r2r_bench_$1_$2.err
done
