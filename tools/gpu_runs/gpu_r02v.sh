set -x
(timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_soak_gpu.py tests/test_host_gpu.py tests/test_errors_gpu.py -x -q 2>&1 | tail -15) > gpurun_out/r02v_pytest.log
tail -4 gpurun_out/r02v_pytest.log
timeout 600 python tools/soak_gpu.py --iters 400 --seed 777 > gpurun_out/r02v_soak.log 2>&1; tail -n 2 gpurun_out/r02v_soak.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02v_launches_c5.csv python bench.py --only c5 --no-e2e --no-cpu --steps 1 --warmup 1 --records 1000000 > gpurun_out/r02v_ncu.log 2>&1
timeout 600 python bench.py --only c5 --no-e2e > gpurun_out/r02v_bench_c5.json 2> gpurun_out/r02v_bench.err; tail -n 2 gpurun_out/r02v_bench.err
python -c "
import json; j=json.load(open('gpurun_out/r02v_bench_c5.json')); c=j['c5']; print('c5 ms', c['ms_per_step'], 'launches', c['launches_per_step'], c['checksum'])"
