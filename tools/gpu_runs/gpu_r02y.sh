set -x
(timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_soak_gpu.py tests/test_host_gpu.py tests/test_errors_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -15) > gpurun_out/r02y_pytest.log
tail -3 gpurun_out/r02y_pytest.log
timeout 600 python tools/soak_gpu.py --iters 300 --seed 4242 > gpurun_out/r02y_soak.log 2>&1; tail -n 1 gpurun_out/r02y_soak.log
timeout 600 python bench.py --only c5 --no-e2e > gpurun_out/r02y_bench_c5.json 2> gpurun_out/r02y_bench.err; tail -n 2 gpurun_out/r02y_bench.err
python -c "
import json; j=json.load(open('gpurun_out/r02y_bench_c5.json')); c=j['c5']; print('c5 ms', c['ms_per_step'], 'launches', c['launches_per_step'], c['checksum']['equal'], c['roofline']['frac'])"
