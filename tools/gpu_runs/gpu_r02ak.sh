set -x
timeout 600 python bench.py --only c4 --no-e2e --no-cpu > gpurun_out/r02ak_c4.json 2> gpurun_out/r02ak.err; tail -2 gpurun_out/r02ak.err
python -c "
import json; j=json.load(open('gpurun_out/r02ak_c4.json')); print(j['c4']['ms_per_step']); print(j['c4']['parts']['nthash'])"
