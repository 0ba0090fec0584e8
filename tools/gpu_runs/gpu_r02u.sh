set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02u_launches_c5.csv python bench.py --only c5 --no-e2e --no-cpu --steps 1 --warmup 1 --records 1000000 > gpurun_out/r02u_ncu.log 2>&1
tail -2 gpurun_out/r02u_ncu.log | cut -c1-300
