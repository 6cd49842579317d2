#!/bin/bash
# round 2, GPU call H: CTA pairs also for the 64-channel layers; full gpu suite with the defaults; smoke; compact ncu evidence
mkdir -p gpurun_out
UB_CASE_TIMEOUT=300 UB_PROBE_OUT=r2h_probe_64.json timeout 1500 python tests/gpu_probe.py conv_fwd_64_64,conv_fwd_big,conv_fwd_small,conv_dgrad_64_64,conv_fwd_folded_64,conv_fwd_folded_cat,conv_fwd_folded_2x2,conv_fwd_bn_64,conv_fwd_bn_2x2,layer_enc1b,layer_dec1a > gpurun_out/r2h_probe_64.log 2>&1; echo "probe 64 rc=$?"
cut -c1-300 gpurun_out/r2h_probe_64.log
for v in 0 1; do
  UB_CONV3_2CTA_64=$v timeout 300 python tools/sustained.py 2 enc1b_fwd enc1b_dgrad > gpurun_out/r2h_sustained_64_$v.jsonl 2> gpurun_out/r2h_sustained_64_$v.err; echo "sustained pairs64=$v rc=$?"; cat gpurun_out/r2h_sustained_64_$v.jsonl
done
for cfg in "1 1" "1 0" "0 0" "1 1"; do
  set -- $cfg
  UB_CONV3_2CTA=$1 UB_CONV3_2CTA_64=$2 timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2h_bench_p$1_$2.json 2> gpurun_out/r2h_bench_p$1_$2.err
  echo "pairs=$1 pairs64=$2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2h_bench_p$1_$2.json'));k=d['kernel_ms_per_step'];print(round(d['ms_per_step'],3),round(d['value'],1),d['clocks']['sm_mhz'],d.get('final_loss'),'fwd',k.get('ub_conv3x3_fwd_bn'),'dgrad',k.get('ub_conv3x3_dgrad'),k.get('ub_conv3x3_dgrad_bnred'))")"
done
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log | cut -c1-600
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2h_smoke.log | cut -c1-600
# compact ncu evidence: --set full on the tensor-core kernels of one forward+backward (report kept in /tmp, only the CSV comes back)
timeout 900 ncu --set full --clock-control none -k regex:'conv3_kernel|conv3_pair_kernel|wgrad_halo_kernel' -s 70 -c 34 -o /tmp/r2h_tensor python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2h_ncu_tensor.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2h_tensor.ncu-rep --page raw --csv > gpurun_out/raw_r2h_tensor.csv 2> /dev/null; wc -c gpurun_out/raw_r2h_tensor.csv
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 620 --csv \
  --log-file gpurun_out/r2h_launches.csv python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2h_ncu.log 2>&1; echo "ncu launches rc=$?"
du -sh gpurun_out
