set -x
timeout 900 python tools/soak_bgzf.py --iters 500 --seed 2024 > gpurun_out/r02ad_soak_bgzf.log 2>&1; tail -n 5 gpurun_out/r02ad_soak_bgzf.log
