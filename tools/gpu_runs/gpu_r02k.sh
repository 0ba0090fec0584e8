set -x
SHB_LIB=seqhasher_b200/_variants/lib_bounds.so timeout 900 python tools/bounds_check.py --iters 300 > gpurun_out/r02k_bounds.log 2>&1
cat gpurun_out/r02k_bounds.log | tail -3
python tools/probe_sha1_variants.py > gpurun_out/r02k_sha1_variants.log 2>&1
cat gpurun_out/r02k_sha1_variants.log
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02k_pytest.log
cat gpurun_out/r02k_pytest.log
timeout 300 python tools/bench_configs.py --only C3 --scale 0.2 --steps 2 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hash_tile_kernel -s 3 -c 1 -o gpurun_out/r02k_c3_tile python tools/bench_configs.py --only C3 --scale 0.2 --steps 2 > gpurun_out/ncu_c3.log 2>&1
timeout 300 python tools/bench_configs.py --only C4 --scale 0.05 --steps 2 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'b3_chunks' -s 3 -c 1 -o gpurun_out/r02k_c4_b3nt python tools/bench_configs.py --only C4 --scale 0.05 --steps 2 > gpurun_out/ncu_c4.log 2>&1
