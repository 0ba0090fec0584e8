set -x
(timeout 900 python -m pytest tests/test_bgzf_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/r02r_pytest.log
cat gpurun_out/r02r_pytest.log
timeout 1200 python tools/bench_bgzf.py --records 1000000 > gpurun_out/r02r_bgzf.json 2> gpurun_out/r02r_bgzf.err; tail -3 gpurun_out/r02r_bgzf.err; cat gpurun_out/r02r_bgzf.json
