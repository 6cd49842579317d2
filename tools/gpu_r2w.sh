#!/bin/bash
# round 2, call W (2 GPUs): sharded tiled inference of the 20000^2 image on the final build
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload config5 --steps 2 --warmup 1 > gpurun_out/r2w_cfg5_n2.json 2> gpurun_out/r2w_cfg5_n2.err; echo "cfg5 n2 rc=$?"
cut -c1-400 gpurun_out/r2w_cfg5_n2.json; tail -2 gpurun_out/r2w_cfg5_n2.err
