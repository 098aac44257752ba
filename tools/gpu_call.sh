mkdir -p gpurun_out
timeout 600 python tools/train_loop.py --strands 5000 --iters 200 --densify-every 20 --profile 2>&1 | tail -1 > gpurun_out/r02_k_loop_500k_profile.json
timeout 900 python tools/train_loop.py --strands 20000 --iters 300 --densify-every 100 2>&1 | tail -1 > gpurun_out/r02_k_config5_mine.json
timeout 1500 python tools/train_loop.py --strands 20000 --iters 300 --densify-every 100 --impl reference 2>&1 | tail -1 > gpurun_out/r02_k_config5_reference.json
cat gpurun_out/r02_k_loop_500k_profile.json gpurun_out/r02_k_config5_mine.json gpurun_out/r02_k_config5_reference.json
