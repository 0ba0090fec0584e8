set -x
(timeout 900 python -m pytest tests/test_bgzf_gpu.py -x -q 2>&1 | tail -25) > gpurun_out/r02n_pytest.log
cat gpurun_out/r02n_pytest.log
