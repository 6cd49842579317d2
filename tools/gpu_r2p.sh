#!/bin/bash
# round 2, call P: rows kernel on by default (real default this time), vectorised RED loop, RED variant of the rows kernel (UB_FUSE_RED64)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "(conv_fwd or conv_dgrad or deconv_ or layer_ or rows_ or fold) and not wgrad and not botb" > gpurun_out/r2p_pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2p_pytest_kernels.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2p_smoke.log | cut -c1-300
UB_CASE_TIMEOUT=200 UB_PROBE_OUT=r2p_probe.json timeout 600 python tests/gpu_probe.py red64_wellcond,golden_c1_k2_bf16,inference_bf16 > gpurun_out/r2p_probe.log 2>&1; echo "probe rc=$?"; cut -c1-300 gpurun_out/r2p_probe.log
timeout 120 python tools/sustained.py 1.5 enc1b_fwd enc1b_dgrad dec1a_fwd dec1a_dgrad enc2a_dgrad dec2a_fwd > gpurun_out/r2p_sustained.jsonl 2> gpurun_out/r2p_sustained.err; echo "sustained rc=$?"; cat gpurun_out/r2p_sustained.jsonl; tail -3 gpurun_out/r2p_sustained.err
for v in 0 1 0 1; do
  UB_FUSE_RED64=$v timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2p_bench_red$v.json 2> gpurun_out/r2p_bench_red$v.err
  echo "red64=$v rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2p_bench_red$v.json'));k=d['kernel_ms_per_step'];print(round(d['ms_per_step'],3),round(d['value'],1),d['clocks']['sm_mhz'],'fwd',k.get('ub_conv3x3_fwd_bn'),'dgrad',k.get('ub_conv3x3_dgrad'),'bnred',k.get('ub_conv3x3_dgrad_bnred'),'reduce',k.get('ub_bn_bwd_reduce'),'loss',d['final_loss'],'roof',round(d['roofline']['frac'],3))")"
done
timeout 300 python bench.py --workload config5 --steps 2 --warmup 1 > gpurun_out/r2p_cfg5.json 2> gpurun_out/r2p_cfg5.err; echo "cfg5 rc=$?"; cut -c1-200 gpurun_out/r2p_cfg5.json
