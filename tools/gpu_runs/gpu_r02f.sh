set -x
timeout 600 python tools/ab_c4.py seqhasher_b200/libseqhash_b200.so seqhasher_b200/_variants/lib_nt3.so seqhasher_b200/_variants/lib_nt5.so seqhasher_b200/_variants/lib_nt6.so > gpurun_out/r02f_ab_c4.log 2>&1
cat gpurun_out/r02f_ab_c4.log
(timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_soak_gpu.py tests/test_errors_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/r02f_pytest.log
cat gpurun_out/r02f_pytest.log
timeout 300 python tools/bench_configs.py --only C3 --scale 0.2 --steps 2 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hash_tile_kernel -s 3 -c 1 -o gpurun_out/r02f_c3_tile python tools/bench_configs.py --only C3 --scale 0.2 --steps 2 > gpurun_out/ncu_c3.log 2>&1
