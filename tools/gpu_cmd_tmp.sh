python tools/mma_rate.py > gpurun_out/mma_rate_r01g.jsonl 2>&1; tail -45 gpurun_out/mma_rate_r01g.jsonl
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_r01g.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv3_kernel -s 0 -c 2 -f -o gpurun_out/prof_r01g_conv3_first $CMD > gpurun_out/ncu_r01g.log 2>&1
ls -la gpurun_out
