#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
GH_DENSIFY_PROFILE=1 timeout 300 $TR --master-port 29512 tools/train_loop.py --strands 20000 --iters 300 2> gpurun_out/c5_8gpu.err | tail -1 > gpurun_out/r02_final_config5_mine_8gpu.json
grep "densify\]" gpurun_out/c5_8gpu.err | sort | uniq -c | sort -k2 | head -40
cut -c1-600 gpurun_out/r02_final_config5_mine_8gpu.json
