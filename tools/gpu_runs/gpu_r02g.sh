set -x
timeout 600 python tools/ab_c3.py seqhasher_b200/libseqhash_b200.so seqhasher_b200/_variants/lib_skip0.so > gpurun_out/r02g_ab.log 2>&1
cat gpurun_out/r02g_ab.log
(timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_soak_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/r02g_pytest.log
cat gpurun_out/r02g_pytest.log
timeout 300 python tools/bench_configs.py --only C4split --algos nthash,xxh3 --scale 0.05 --steps 2 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'nthash_warp|xxh3_warp' -s 4 -c 2 -o gpurun_out/r02g_c4_warp python tools/bench_configs.py --only C4split --algos nthash,xxh3 --scale 0.05 --steps 2 > gpurun_out/ncu_c4.log 2>&1
