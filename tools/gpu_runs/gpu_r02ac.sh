set -x
for v in upbase uphi upbase uphi; do
  SHB_LIB=$PWD/seqhasher_b200/_variants/lib_$v.so timeout 600 python bench.py --only c2 --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/r02ac_$v.json 2> gpurun_out/r02ac_$v.err
  python -c "
import json; j=json.load(open('gpurun_out/r02ac_$v.json')); print('$v', 'sha1 ms', j['ms_per_step'], 'xxh3 ms', j['xxh3']['ms_per_step'])"
done
SHB_LIB=$PWD/seqhasher_b200/_variants/lib_uphi.so timeout 600 python bench.py --only c3 --no-e2e --no-cpu > gpurun_out/r02ac_c3_uphi.json 2>/dev/null; python -c "
import json; j=json.load(open('gpurun_out/r02ac_c3_uphi.json')); print('uphi c3', j['c3']['ms_per_step'], j['c3']['checksum']['equal'], j['c3']['softmasked_2pct']['ms_per_step'])"
