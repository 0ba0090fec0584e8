set -x
timeout 900 python tools/soak_gpu.py --iters 1500 --seed 77 > gpurun_out/r02l_soak.log 2>&1; tail -2 gpurun_out/r02l_soak.log
timeout 600 python tools/soak_gpu.py --iters 150 --seed 78 --big > gpurun_out/r02l_soak_big.log 2>&1; tail -2 gpurun_out/r02l_soak_big.log
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
tail -2 gpurun_out/ncu_bench.err
wc -l gpurun_out/r02_launches_bench.csv
