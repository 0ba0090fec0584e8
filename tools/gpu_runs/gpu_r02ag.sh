set -x
timeout 900 python bench.py --only c5 --no-cpu > gpurun_out/r02ag_bench_c5.json 2> gpurun_out/r02ag_bench.err; tail -n 2 gpurun_out/r02ag_bench.err
python - <<'P'
import json
j=json.load(open('gpurun_out/r02ag_bench_c5.json')); g=j['c5']['e2e_gzip_cli']
print({k:v for k,v in g.items() if k in ('records','gzip_s','plain_s','same_output_as_plain','text_bytes')})
print(g['bgzf']); print(g['bgzf_in_process'])
for k,v in g.items():
    if 'phases' in k: print(k, v)
P
