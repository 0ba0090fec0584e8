set -x
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02af_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02af_ncu.log 2>&1
grep -c . gpurun_out/r02af_launches_bench.csv; tail -c 300 gpurun_out/r02af_ncu.log
