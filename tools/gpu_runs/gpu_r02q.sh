set -x
(timeout 900 python -m pytest tests/test_bgzf_gpu.py -x -q -k "cli_bgzf_equals" 2>&1 | tail -60) > gpurun_out/r02q_pytest.log
grep -n "Error\|assert\|env\|stderr" gpurun_out/r02q_pytest.log | head -20
