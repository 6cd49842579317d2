#!/bin/bash
# A/B of the level-1 conv pipeline shapes (UB_CONV3_64 / UB_CONV3_64B), per-layer microbench
mkdir -p gpurun_out; out=gpurun_out/layers_micro_v5.jsonl; : > $out
for v in 0 1 2 3 4 5; do echo "{\"UB_CONV3_64\": $v}" >> $out; UB_CONV3_64=$v BL_MODES=fwd,dgrad timeout 120 python tools/bench_layers.py enc1b >> $out 2>&1; done
for v in 0 1 2; do echo "{\"UB_CONV3_64B\": $v}" >> $out; UB_CONV3_64B=$v BL_MODES=fwd timeout 120 python tools/bench_layers.py dec1a >> $out 2>&1; done
cat $out
