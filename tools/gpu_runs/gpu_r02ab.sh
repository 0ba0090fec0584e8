# usage: bash tools/gpu_runs/gpu_r02ab.sh N   (under gpurun --gpus N): bench line + the multi-GPU CLI tests
N=$1
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02ab_bench_n$N.json 2> gpurun_out/r02ab_bench_n$N.err
tail -3 gpurun_out/r02ab_bench_n$N.err
python - <<P
import json
j=json.load(open('gpurun_out/r02ab_bench_n$N.json'))
print('c2', j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'])
for k in ('c3','c4','c5'):
    c=j[k]; print(k, c['ms_per_step'], c['checksum']['equal'], c.get('e2e',{}).get('value'))
P
(timeout 600 python -m pytest tests/test_cli_binary.py -x -q -m gpu -k "two_gpus or small_chunks" 2>&1 | tail -3) > gpurun_out/r02ab_pytest_multi_n$N.log
cat gpurun_out/r02ab_pytest_multi_n$N.log
