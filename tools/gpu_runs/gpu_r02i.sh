set -x
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02i_pytest.log
cat gpurun_out/r02i_pytest.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/r02i_smoke.log
cat gpurun_out/r02i_smoke.log
timeout 1200 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
tail -5 gpurun_out/r02i_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02i_bench_ref.json 2>> gpurun_out/r02i_bench.err
