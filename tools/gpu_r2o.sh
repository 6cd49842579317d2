#!/bin/bash
# round 2, call O: vectorised statistics read-back in the shared epilogue, row-streaming kernel on by default (ragged last strip allowed)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "(conv_fwd or conv_dgrad or deconv_ or layer_enc or layer_dec1a or rows_ or fold) and not wgrad" > gpurun_out/r2o_pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2o_pytest_kernels.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2o_smoke.log | cut -c1-600
timeout 120 python tools/sustained.py 1.5 enc1b_fwd enc1b_dgrad dec1a_fwd enc2a_fwd enc3b_fwd > gpurun_out/r2o_sustained.jsonl 2> gpurun_out/r2o_sustained.err; echo "sustained rc=$?"; cat gpurun_out/r2o_sustained.jsonl; tail -3 gpurun_out/r2o_sustained.err
UB_CONV3_ROWS=0 timeout 120 python tools/sustained.py 1.5 enc1b_fwd enc2a_fwd > gpurun_out/r2o_sustained_rows0.jsonl 2> gpurun_out/r2o_sustained_rows0.err; cat gpurun_out/r2o_sustained_rows0.jsonl
for v in 0 2 0 2; do
  UB_CONV3_ROWS=$v timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2o_bench_rows$v.json 2> gpurun_out/r2o_bench_rows$v.err
  echo "rows=$v rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2o_bench_rows$v.json'));k=d['kernel_ms_per_step'];print(round(d['ms_per_step'],3),round(d['value'],1),d['clocks']['sm_mhz'],'fwd',k.get('ub_conv3x3_fwd_bn'),'dgrad',k.get('ub_conv3x3_dgrad'),'deconv_fwd',k.get('ub_deconv2x2_fwd_bn'),'loss',d['final_loss'],'roof',round(d['roofline']['frac'],3))")"
done
timeout 300 python bench.py --workload config5 --steps 2 --warmup 1 > gpurun_out/r2o_cfg5.json 2> gpurun_out/r2o_cfg5.err; echo "cfg5 rc=$?"; cut -c1-400 gpurun_out/r2o_cfg5.json
