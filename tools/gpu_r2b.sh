#!/bin/bash
# round 2, GPU call B: full gpu suite with the promoted cases, numbers for the trained-weights / tiled-inference parity cases,
# A/B of the BatchNorm fold (border sums fixed) x fused elementwise reductions, config4 / config5 bench lines, whole-step DRAM traffic.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2b_pytest.log | cut -c1-1200
UB_VERBOSE=1 UB_CASE_TIMEOUT=900 UB_PROBE_OUT=r2b_probe.json timeout 1500 python tests/gpu_probe.py probe_trained,tiled_inference_bf16,banded_inference,conv_first_tiles,inference_bf16 > gpurun_out/r2b_probe.log 2>&1; echo "probe rc=$?"
cut -c1-1800 gpurun_out/r2b_probe.log
for cfg in "0 0" "1 0" "1 1" "0 1"; do
  set -- $cfg
  UB_FOLD_BN=$1 UB_FUSE_EW=$2 timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2b_bench_fold$1_ew$2.json 2> gpurun_out/r2b_bench_fold$1_ew$2.err
  echo "fold=$1 ew=$2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2b_bench_fold$1_ew$2.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks']['sm_mhz'],d['clocks']['samples'],d.get('final_loss'))")"
done
timeout 300 python bench.py --workload config4 --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2b_bench_cfg4.json 2> gpurun_out/r2b_bench_cfg4.err; echo "cfg4 rc=$? $(cut -c1-300 gpurun_out/r2b_bench_cfg4.json)"
timeout 600 python bench.py --workload config5 --steps 3 > gpurun_out/r2b_bench_cfg5.json 2> gpurun_out/r2b_bench_cfg5.err; echo "cfg5 rc=$? $(cut -c1-400 gpurun_out/r2b_bench_cfg5.json)"; tail -3 gpurun_out/r2b_bench_cfg5.err
# whole-step DRAM traffic + launch list (serialised, cold cache: shares and bytes only)
UB_FOLD_BN=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/r2b_launches_fold1.csv python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
