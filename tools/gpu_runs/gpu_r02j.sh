set -x
(timeout 900 python -m pytest tests/test_cli_binary.py tests/test_host_gpu.py -x -q -m gpu 2>&1 | tail -5) > gpurun_out/r02j_pytest.log
cat gpurun_out/r02j_pytest.log
python tools/bench_cli_multi.py --gb 8 --gpus 1,2 > gpurun_out/r02j_cli_multi.log 2>&1
cat gpurun_out/r02j_cli_multi.log
SHB_NO_RANGES=1 python tools/bench_cli_multi.py --gb 8 --gpus 1 > gpurun_out/r02j_cli_stream.log 2>&1
cat gpurun_out/r02j_cli_stream.log
