#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/gpu_probe.py augment,checkpoint,reader_augmented,train_driver,tiled_inference_bf16 > gpurun_out/probe_new.log 2>&1; echo "probe rc=$?"
cat gpurun_out/probe_new.log | cut -c1-900
