#!/bin/bash
# round 2, GPU call E: BatchNorm sums from wgrad (parity + A/B), head-backward loads in flight, ncu --set full of the tensor-core kernels
mkdir -p gpurun_out
UB_CASE_TIMEOUT=900 UB_PROBE_OUT=r2e_probe.json timeout 2400 python tests/gpu_probe.py bn_sums_,live_bf16,wellcond_bf16,golden_c1_k2_bf16,golden_c3_k8_bf16,trained_bf16,config1_refdata,head_k,sharded_inference,checkpoint_roundtrip > gpurun_out/r2e_probe.log 2>&1; echo "probe rc=$?"
cut -c1-500 gpurun_out/r2e_probe.log
for cfg in "1 4" "0 4" "1 2" "1 4"; do
  set -- $cfg
  UB_BN_ALGEBRA=$1 UB_HEAD_PPI=$2 timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2e_bench_alg$1_ppi$2.json 2> gpurun_out/r2e_bench_alg$1_ppi$2.err
  echo "alg=$1 ppi=$2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2e_bench_alg$1_ppi$2.json'));k=d['kernel_ms_per_step'];print(round(d['ms_per_step'],3),round(d['value'],1),round(d['e2e']['value'],1),d['clocks']['sm_mhz'],d['gpu_launches']//30,d.get('final_loss'),'head_bwd',k.get('ub_head_bwd_apply'),'bnred',k.get('ub_bn_bwd_reduce'),'dgrad_red',k.get('ub_conv3x3_dgrad_bnred'))")"
done
UB_OVERLAP_WGRAD=0 timeout 300 python bench.py --no-cpu-baseline --steps 30 --warmup 6 > gpurun_out/r2e_bench_noovl.json 2> gpurun_out/r2e_bench_noovl.err
echo "no-overlap rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2e_bench_noovl.json'));print(round(d['ms_per_step'],3),round(d['value'],1))")"
timeout 900 ncu --set full --clock-control none -k regex:'conv3_kernel|wgrad_halo_kernel|igemm_' -s 130 -c 60 -o gpurun_out/r2e_tensor python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2e_ncu_tensor.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2e_tensor.ncu-rep --page raw --csv > gpurun_out/raw_r2e_tensor.csv 2> /dev/null; wc -l gpurun_out/raw_r2e_tensor.csv
UB_BN_ALGEBRA=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 640 --csv \
  --log-file gpurun_out/r2e_launches.csv python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2e_ncu.log 2>&1; echo "ncu launches rc=$?"
