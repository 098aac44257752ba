#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_final2_pytest_gpu.txt
python bench.py --no-cpu-baseline 2> gpurun_out/r02_final2_bench_native.err | tail -1 > gpurun_out/r02_final2_bench_native.json
python tools/train_loop.py --strands 20000 --iters 300 2>/dev/null | tail -1 > gpurun_out/r02_final2_config5_mine.json
cat gpurun_out/r02_final2_pytest_gpu.txt
grep "bench\]" gpurun_out/r02_final2_bench_native.err | tail -8
cut -c1-330 gpurun_out/r02_final2_config5_mine.json
