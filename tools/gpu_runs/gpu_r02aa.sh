set -x
timeout 1500 python bench.py > gpurun_out/r02aa_bench.json 2> gpurun_out/r02aa_bench.err; tail -n 3 gpurun_out/r02aa_bench.err
python - <<'P'
import json
j=json.load(open('gpurun_out/r02aa_bench.json'))
print('c2', j['ms_per_step'], j['roofline']['frac'], 'e2e', j['e2e']['value'], 'xxh3', j['xxh3']['ms_per_step'], j['xxh3']['roofline']['frac'])
for k in ('c3','c4','c5'):
    c=j[k]; print(k, c['ms_per_step'], c['roofline']['frac'], c['checksum']['equal'], c.get('e2e',{}).get('value'))
print('c4 parts', {k:(v['ms_per_step'], v.get('roofline',{}).get('frac')) for k,v in j['c4']['parts'].items()})
g=j['c5']['e2e_gzip_cli']
print('c5 cli', {k:v for k,v in g.items() if k in ('gzip_s','plain_s','same_output_as_plain')})
print('bgzf', g.get('bgzf')); print('bgzf_in_process', g.get('bgzf_in_process'))
P
