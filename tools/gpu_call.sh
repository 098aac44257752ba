#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tools/train_loop.py --strands 20000 --iters 300 --profile 2>/dev/null | tail -1 > gpurun_out/prof_config5_2gpu.json
timeout 300 python tools/train_loop.py --strands 20000 --iters 300 --profile 2>/dev/null | tail -1 > gpurun_out/prof_config5_1gpu.json
timeout 300 $TR --master-port 29513 tools/train_loop.py --strands 20000 --iters 300 2>/dev/null | tail -1 > gpurun_out/config5_2gpu.json
cut -c1-1500 gpurun_out/prof_config5_2gpu.json gpurun_out/prof_config5_1gpu.json gpurun_out/config5_2gpu.json
