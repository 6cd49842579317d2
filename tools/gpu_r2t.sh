#!/bin/bash
# round 2, call T: the whole GPU suite without -x (call S stopped at the first failure)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2t_pytest.log | cut -c1-900
