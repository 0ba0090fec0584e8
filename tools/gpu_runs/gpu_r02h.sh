set -x
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_soak_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/r02h_pytest.log
cat gpurun_out/r02h_pytest.log
timeout 600 python tools/ab_c4.py seqhasher_b200/libseqhash_b200.so > gpurun_out/r02h_ab_c4.log 2>&1
cat gpurun_out/r02h_ab_c4.log
timeout 600 python tools/ab_c3.py seqhasher_b200/libseqhash_b200.so > gpurun_out/r02h_ab_c3.log 2>&1
cat gpurun_out/r02h_ab_c3.log
