set -x
mkdir -p gpurun_out
nvidia-smi -L
nproc; free -g | head -2
python tools/probe_int64.py > gpurun_out/r02a_int64.log 2>&1
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02a_pytest.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5) > gpurun_out/r02a_smoke.log
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_bench.err
timeout 300 python tools/bench_configs.py --only C3 --scale 0.2 --steps 2 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hash_tile_kernel -s 3 -c 1 -o gpurun_out/r02a_c3_tile python tools/bench_configs.py --only C3 --scale 0.2 --steps 2 > gpurun_out/ncu_c3.log 2>&1
timeout 300 python tools/bench_configs.py --only C4split --algos nthash,xxh3 --scale 0.05 --steps 2 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'nthash_warp|xxh3_warp' -s 4 -c 2 -o gpurun_out/r02a_c4_warp python tools/bench_configs.py --only C4split --algos nthash,xxh3 --scale 0.05 --steps 2 > gpurun_out/ncu_c4.log 2>&1
ls -la gpurun_out
