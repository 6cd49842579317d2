CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
for k in conv_first_wgrad_strip conv_first_fwd_strip pool_bwd_add igemm_wgrad_kernel head_bwd_apply igemm_fwd_kernel; do
  $CMD > gpurun_out/plain_r01j.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 4 -f -o /tmp/prof_$k $CMD > gpurun_out/ncu_r01j_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/raw_r01j_$k.csv 2>/dev/null
  ncu -i /tmp/prof_$k.ncu-rep --page source --csv --print-source sass > gpurun_out/src_r01j_$k.csv 2>/dev/null
done
du -sh gpurun_out
