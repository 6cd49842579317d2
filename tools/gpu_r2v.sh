#!/bin/bash
# round 2, call V (2 GPUs): the data-parallel GPU test that a 1-GPU box skips
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_dp_gpu.py -q -m gpu > gpurun_out/r2v_pytest_dp.log 2>&1; echo "pytest dp rc=$?"; tail -4 gpurun_out/r2v_pytest_dp.log | cut -c1-400
