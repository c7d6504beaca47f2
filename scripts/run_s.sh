# last run of the round: GPU suite on the final tree
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_s_pytest.log
cat gpurun_out/r02_s_pytest.log
