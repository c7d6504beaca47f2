timeout 45 python -m pytest tests/test_host_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_u_pytest.log
cat gpurun_out/r02_u_pytest.log
