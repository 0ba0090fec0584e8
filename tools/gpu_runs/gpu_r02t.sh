set -x
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -x -q -k "long or nthash or c4 or C4" 2>&1 | tail -15) > gpurun_out/r02t_pytest.log
tail -4 gpurun_out/r02t_pytest.log
timeout 600 python tools/soak_gpu.py --iters 300 --seed 123 > gpurun_out/r02t_soak.log 2>&1; tail -n 2 gpurun_out/r02t_soak.log
timeout 600 python tools/ab_c4.py seqhasher_b200/_variants/lib_ntlut.so seqhasher_b200/_variants/lib_ntq3.so seqhasher_b200/_variants/lib_ntq4.so seqhasher_b200/_variants/lib_ntq5.so seqhasher_b200/_variants/lib_ntq6.so > gpurun_out/r02t_ab_c4.log 2>&1; cat gpurun_out/r02t_ab_c4.log
