#!/bin/bash
# round 2, final call: the whole GPU suite, smoke, the bench lines of the three workloads, ncu evidence of the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2s_pytest.log | cut -c1-600
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2s_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 6 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/r2s_bench.json
timeout 300 python bench.py --workload config4 --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2s_cfg4.json 2> gpurun_out/r2s_cfg4.err; echo "cfg4 rc=$?"; cut -c1-300 gpurun_out/r2s_cfg4.json
timeout 300 python bench.py --workload config5 --steps 3 --warmup 1 > gpurun_out/r2s_cfg5.json 2> gpurun_out/r2s_cfg5.err; echo "cfg5 rc=$?"; cut -c1-300 gpurun_out/r2s_cfg5.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 620 --csv \
  --log-file gpurun_out/launches_r2s.csv python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2s_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'conv3_rows_kernel' -s 14 -c 10 -o /tmp/r2s_rows python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2s_ncu_rows.log 2>&1; echo "ncu rows rc=$?"
ncu -i /tmp/r2s_rows.ncu-rep --page raw --csv > gpurun_out/raw_r2s_rows.csv 2> /dev/null; wc -c gpurun_out/raw_r2s_rows.csv
du -sh gpurun_out
