set -x
(timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r02s_pytest.log
tail -5 gpurun_out/r02s_pytest.log
