mkdir -p gpurun_out
TAG=r02_j
timeout 1500 python -m pytest tests/test_gpu_projection.py -m gpu -q 2>&1 | tail -6
timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
grep "bench\]" gpurun_out/${TAG}_bench.err | tail -6
