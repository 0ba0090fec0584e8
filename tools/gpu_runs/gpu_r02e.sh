set -x
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02e_pytest.log
cat gpurun_out/r02e_pytest.log
timeout 600 python tools/ab_c3.py seqhasher_b200/libseqhash_b200.so seqhasher_b200/_variants/lib_mad0.so > gpurun_out/r02e_ab.log 2>&1
cat gpurun_out/r02e_ab.log
timeout 600 python tools/bench_configs.py --only C4split --algos nthash,xxh3,blake3 --scale 0.2 --steps 3 > gpurun_out/r02e_c4.log 2>&1
cat gpurun_out/r02e_c4.log | cut -c1-400
