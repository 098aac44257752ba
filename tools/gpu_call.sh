mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_d_pytest.txt
cat gpurun_out/r02_d_pytest.txt
timeout 600 python bench.py --steps 30 --warmup 5 2> gpurun_out/r02_d_bench.err | tail -1 > gpurun_out/r02_d_bench.json
grep "bench\]" gpurun_out/r02_d_bench.err | tail -14
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2> gpurun_out/r02_d_bench_ref.err | tail -1 > gpurun_out/r02_d_bench_ref.json
grep "bench\]" gpurun_out/r02_d_bench_ref.err | tail -8
