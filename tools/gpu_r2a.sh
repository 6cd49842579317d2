#!/bin/bash
# round 2, GPU call A: full gpu test suite (incl. the config-2 layer-shape cases), bf16 parity measurements at well-conditioned
# shapes, the pending (never-run) cases, smoke, and the A/B of the BatchNorm fold.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/r2a_gpu.txt
nproc >> gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log | cut -c1-600
UB_VERBOSE=1 UB_CASE_TIMEOUT=1500 UB_PROBE_OUT=r2a_probe_parity.json timeout 3000 python tests/gpu_probe.py --probe > gpurun_out/r2a_probe_parity.log 2>&1; echo "probe rc=$?"
cut -c1-1500 gpurun_out/r2a_probe_parity.log
UB_PROBE_OUT=r2a_probe_pending.json timeout 1500 python tests/gpu_probe.py --pending > gpurun_out/r2a_probe_pending.log 2>&1; echo "pending rc=$?"
cut -c1-500 gpurun_out/r2a_probe_pending.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2a_smoke.log | cut -c1-1500
for f in 0 1; do
  UB_FOLD_BN=$f timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2a_bench_fold$f.json 2> gpurun_out/r2a_bench_fold$f.err
  echo "fold=$f rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2a_bench_fold$f.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks']['sm_mhz'],d.get('final_loss'))")"
done
