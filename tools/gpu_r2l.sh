#!/bin/bash
# round 2, GPU call L: final verification of the committed build (what the driver runs at round end) + sanitizer sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l_pytest.log | cut -c1-400
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2l_smoke.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2l_bench.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks'],d['roofline']['achieved'],d['roofline']['frac'],d['cpu_baseline'])")"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err; echo "ref rc=$? $(cut -c1-300 gpurun_out/r2l_bench_ref.json)"
timeout 400 python bench.py --workload config5 --steps 3 > gpurun_out/r2l_cfg5.json 2> gpurun_out/r2l_cfg5.err; echo "cfg5 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2l_cfg5.json'));print(round(d['value'],1),round(d['e2e']['value'],1),d['roofline']['achieved'],d['mask_crc'])")"
timeout 400 python bench.py --workload config4 --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2l_cfg4.json 2> gpurun_out/r2l_cfg4.err; echo "cfg4 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2l_cfg4.json'));print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1))")"
timeout 1500 bash tools/sanitize.sh > gpurun_out/r2l_sanitize.log 2>&1; echo "sanitize rc=$?"; cat gpurun_out/r2l_sanitize.log | cut -c1-200
