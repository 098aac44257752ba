mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29631 tools/allreduce_case.py 500000 --sweep > gpurun_out/r02_m_allreduce8.txt 2> gpurun_out/r02_m_allreduce8.err
grep -v "^{" gpurun_out/r02_m_allreduce8.txt; tail -2 gpurun_out/r02_m_allreduce8.err
timeout 600 $TR --master-port 29632 bench.py --gpus 8 --steps 30 --warmup 5 2> gpurun_out/r02_m_bench8.err | tail -1 > gpurun_out/r02_m_bench8.json
grep "bench\] rank 0\|bench\] mine\|bench\] e2e\|bench\] train" gpurun_out/r02_m_bench8.err | sort | uniq | tail -8
timeout 900 $TR --master-port 29633 tools/train_loop.py --strands 20000 --iters 300 --densify-every 100 2> gpurun_out/r02_m_config5_8gpu.err | tail -1 > gpurun_out/r02_m_config5_8gpu.json
cat gpurun_out/r02_m_config5_8gpu.json; tail -2 gpurun_out/r02_m_config5_8gpu.err
timeout 900 $TR --master-port 29634 tools/train_loop.py --strands 5000 --iters 300 --no-densify 2> gpurun_out/r02_m_config4_8gpu.err | tail -1 > gpurun_out/r02_m_config4_8gpu.json
cat gpurun_out/r02_m_config4_8gpu.json
