timeout 200 python -m pytest tests/test_packed_gpu.py -m gpu -x -q -k "mixed_packed" 2>&1 | tail -30 > gpurun_out/r02_t_pytest.log
cat gpurun_out/r02_t_pytest.log
