#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tests/gpu_probe.py augment,reader_augmented,train_driver > gpurun_out/probe_final.log 2>&1; echo "probe rc=$?"
cut -c1-420 gpurun_out/probe_final.log
timeout 30 python tools/bench_input.py > gpurun_out/input_r01y2.json 2> gpurun_out/input_r01y2.err; echo "input rc=$?"; cut -c1-330 gpurun_out/input_r01y2.json
