set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hash_tile_kernel -c 4 -o gpurun_out/r02aj_c2_tile python bench.py --only c2 --no-e2e --no-cpu --steps 1 --warmup 1 > gpurun_out/r02aj_ncu.log 2>&1; tail -1 gpurun_out/r02aj_ncu.log | cut -c1-200
