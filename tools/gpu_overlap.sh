#!/bin/bash
# Side-stream weight-gradient experiment: parity suite with the overlap on, then bench variants (device-resident + e2e lines).
mkdir -p gpurun_out
UB_OVERLAP_WGRAD=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ovl.log 2>&1; echo "pytest(overlap) rc=$?" | tee -a gpurun_out/pytest_gpu_ovl.log
tail -3 gpurun_out/pytest_gpu_ovl.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/bench_ovl_$tag.json 2> gpurun_out/bench_ovl_$tag.err; echo "$tag rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_ovl_$tag.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks'])")"; }
run base UB_OVERLAP_WGRAD=0
run ovl UB_OVERLAP_WGRAD=1
run ovl_bn2 UB_OVERLAP_WGRAD=1 UB_BN_BWD_CTAS=2
run ovl_p2 UB_OVERLAP_WGRAD=1 UB_WGRAD_P_STAGES=2
run ovl_bn2_p2 UB_OVERLAP_WGRAD=1 UB_BN_BWD_CTAS=2 UB_WGRAD_P_STAGES=2
run base2 UB_OVERLAP_WGRAD=0
