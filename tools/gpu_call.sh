mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_c_pytest.txt
cat gpurun_out/r02_c_pytest.txt
