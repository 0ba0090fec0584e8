set -x
(timeout 900 python -m pytest tests/test_bgzf_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/r02p_pytest.log
cat gpurun_out/r02p_pytest.log
timeout 900 python tools/bench_bgzf.py > gpurun_out/r02p_bgzf.json 2> gpurun_out/r02p_bgzf.err; tail -3 gpurun_out/r02p_bgzf.err; cat gpurun_out/r02p_bgzf.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bgzf_inflate -c 1 -o gpurun_out/r02p_bgzf_inflate python tools/bench_bgzf.py --records 100000 --no-cli --reps 1 > gpurun_out/r02p_ncu.log 2>&1; tail -3 gpurun_out/r02p_ncu.log
