#!/bin/bash
mkdir -p gpurun_out
timeout 80 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_final3_pytest_gpu.txt
cat gpurun_out/r02_final3_pytest_gpu.txt
