set -x
nvidia-smi -L | head -3
(timeout 900 python -m pytest tests/test_bgzf_gpu.py tests/test_cli_binary.py -x -q -m gpu -k "two_gpus or refused" 2>&1 | tail -6) > gpurun_out/r02ai_pytest.log
cat gpurun_out/r02ai_pytest.log
