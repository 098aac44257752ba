#!/bin/bash
mkdir -p gpurun_out
python tools/bench_variants.py run > gpurun_out/fwd_variants.txt 2>&1
python tools/bench_variants.py run --strands 20000 >> gpurun_out/fwd_variants.txt 2>&1
grep -v "^\[" gpurun_out/fwd_variants.txt
