#!/bin/bash
# One gpurun call: parity tests, bench, ncu launch list, ncu --set full captures of the dominant kernels.
# usage: tools/gpu_round.sh <tag> [kernel regexes for --set full ...]
# gpurun_out/ must stay under 64 MiB: the .ncu-rep of each capture is exported to CSV (raw page) on the box and only
# a short one (-c 2) is kept per kernel for the source page.
tag=${1:-r01}; shift
kernels=${@:-"conv3_kernel wgrad_halo_kernel bn_bwd_apply"}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log
tail -5 gpurun_out/pytest_gpu_$tag.log
python bench.py --layers gpurun_out/layers_$tag.json > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "reference arm rc=$?"; cat gpurun_out/bench_ref_$tag.json
python bench.py --workload config4 --no-cpu-baseline > gpurun_out/bench_cfg4_$tag.json 2> gpurun_out/bench_cfg4_$tag.err; echo "config4 rc=$?"; cut -c1-300 gpurun_out/bench_cfg4_$tag.json
python tools/bench_infer.py --size 20000 > gpurun_out/infer_20000_$tag.json 2> gpurun_out/infer_20000_$tag.err; echo "infer rc=$?"; cat gpurun_out/infer_20000_$tag.json
# the ncu passes use the serial schedule (UB_OVERLAP_WGRAD=0): ncu serialises kernels anyway, and -k/-c launch counting stays per step
CMD="env UB_OVERLAP_WGRAD=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_$tag.log 2>&1
for k in $kernels; do
  $CMD > gpurun_out/plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none -k regex:$k -s 0 -c $( [[ $k == bn_* ]] && echo 6 || echo 17 ) -f -o /tmp/prof_$k $CMD > gpurun_out/ncu_${tag}_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/raw_${tag}_$k.csv 2>/dev/null
  rm -f /tmp/prof_$k.ncu-rep
done
# source-level report (small): the first two conv3 launches of a step = the level-1 64->64 forward (enc1b) and enc2a
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3_kernel -s 0 -c 1 -f -o gpurun_out/prof_${tag}_conv3_enc1b $CMD > /dev/null 2>&1
ncu -i gpurun_out/prof_${tag}_conv3_enc1b.ncu-rep --page source --csv --print-source sass > gpurun_out/src_${tag}_conv3_enc1b.csv 2>/dev/null
du -sh gpurun_out; ls gpurun_out | grep $tag
