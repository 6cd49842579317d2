#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tests/gpu_probe.py zscore,augment_noise,augment_c1,reader_augmented,tiled_inference_bf16 > gpurun_out/probe_final2.log 2>&1; echo "probe rc=$?"
cut -c1-260 gpurun_out/probe_final2.log
timeout 25 python tools/bench_input.py > gpurun_out/input_r01y3.json 2> gpurun_out/input_r01y3.err; echo "input rc=$?"; cut -c1-330 gpurun_out/input_r01y3.json
