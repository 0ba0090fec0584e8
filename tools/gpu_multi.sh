#!/bin/bash
# usage: bash tools/gpu_multi.sh N   (under gpurun --gpus N)
N=$1
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -9
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
nproc; free -g | head -2; lscpu | grep -E "NUMA|Socket|Model name" 
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_roof.py > gpurun_out/h2d_roof_n$N.log 2>&1
grep aggregate gpurun_out/h2d_roof_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err
python tools/bench_cli_multi.py --gb 8 --gpus 1,2,4,8 > gpurun_out/cli_multi_n$N.log 2>&1
cat gpurun_out/cli_multi_n$N.log
(timeout 600 python -m pytest tests/test_cli_binary.py -x -q -m gpu -k "two_gpus or small_chunks" 2>&1 | tail -3) > gpurun_out/pytest_multi_n$N.log
cat gpurun_out/pytest_multi_n$N.log
