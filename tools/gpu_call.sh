#!/bin/bash
# scratch command script for one gpurun call
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
python tools/stage_times.py > gpurun_out/stage_times.txt 2>&1
python tools/stage_times.py --strands 20000 >> gpurun_out/stage_times.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gh_tile_scan" -s 2 -c 1 -o gpurun_out/scan_full \
    python tools/stage_times.py > gpurun_out/scan_full.log 2>&1
cat gpurun_out/pytest_gpu.txt; cat gpurun_out/stage_times.txt | tail -4
