#!/bin/bash
mkdir -p gpurun_out
bash tools/profile_round.sh r02_final > gpurun_out/r02_final_round.log 2>&1
python tools/train_loop.py --strands 20000 --iters 300 2>/dev/null | tail -1 > gpurun_out/r02_final_config5_mine.json
python tools/train_loop.py --strands 5000 --iters 300 2>/dev/null | tail -1 > gpurun_out/r02_final_config4shape_mine.json
tail -30 gpurun_out/r02_final_round.log
cat gpurun_out/r02_final_config5_mine.json gpurun_out/r02_final_config4shape_mine.json | cut -c1-600
