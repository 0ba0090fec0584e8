set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02x_launches_c5.csv python bench.py --only c5 --no-e2e --no-cpu --steps 1 --warmup 1 --records 1000000 > gpurun_out/r02x_ncu.log 2>&1
grep -c . gpurun_out/r02x_launches_c5.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nthash_warp -c 1 -o gpurun_out/r02x_nthash_warp python bench.py --only c4 --no-e2e --no-cpu --steps 1 --warmup 1 --records 1000000 > gpurun_out/r02x_ncu2.log 2>&1; tail -1 gpurun_out/r02x_ncu2.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:compact_kernel -c 1 -o gpurun_out/r02x_compact python bench.py --only c5 --no-e2e --no-cpu --steps 1 --warmup 1 --records 1000000 > gpurun_out/r02x_ncu3.log 2>&1; tail -1 gpurun_out/r02x_ncu3.log | cut -c1-200
timeout 900 python tools/bench_bgzf.py > gpurun_out/r02x_bgzf.json 2> gpurun_out/r02x_bgzf.err; tail -2 gpurun_out/r02x_bgzf.err; python -c "
import json; j=json.load(open('gpurun_out/r02x_bgzf.json')); print(j['cli_hash_sha1']); print(j['cli_all12']); print(j['cli_phases']); print(j['slots_4'])"
