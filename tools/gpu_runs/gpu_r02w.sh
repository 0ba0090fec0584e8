set -x
(timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r02w_pytest.log
tail -3 gpurun_out/r02w_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_smoke.log 2>&1; tail -2 gpurun_out/r02w_smoke.log
timeout 1500 python bench.py > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err; tail -n 3 gpurun_out/r02w_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r02w_bench_ref.json 2> gpurun_out/r02w_bench_ref.err; tail -c 600 gpurun_out/r02w_bench_ref.json
python - <<'P'
import json
j=json.load(open('gpurun_out/r02w_bench.json'))
print('c2', j['ms_per_step'], j['roofline']['frac'], 'e2e', j['e2e']['value'])
for k in ('c3','c4','c5'):
    c=j[k]; print(k, c['ms_per_step'], c['roofline']['frac'], c['checksum']['equal'], c.get('e2e',{}).get('value'))
print('c4 parts', {k:v['ms_per_step'] for k,v in j['c4']['parts'].items()})
print('c5 cli', {k:v for k,v in j['c5']['e2e_gzip_cli'].items() if k in ('gzip_s','plain_s','same_output_as_plain','bgzf')})
P
