mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02_a_pytest.txt
cat gpurun_out/r02_a_pytest.txt
timeout 600 python bench.py --steps 30 --warmup 5 2> gpurun_out/r02_a_bench.err | tail -1 > gpurun_out/r02_a_bench.json
grep "bench\]" gpurun_out/r02_a_bench.err | tail -14
