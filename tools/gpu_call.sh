#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_callers.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
python tools/stage_times.py > gpurun_out/stage_times.txt 2>&1
python tools/stage_times.py --strands 20000 >> gpurun_out/stage_times.txt 2>&1
cat gpurun_out/pytest_gpu.txt; cat gpurun_out/stage_times.txt | tail -4
