set -x
(timeout 1200 python -m pytest tests/test_bgzf_gpu.py tests/test_cli_binary.py tests/test_textproc_gpu.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r02ah_pytest.log
tail -3 gpurun_out/r02ah_pytest.log
timeout 900 python tools/bench_bgzf.py --records 1000000 > gpurun_out/r02ah_bgzf.json 2> gpurun_out/r02ah_bgzf.err; python -c "
import json; j=json.load(open('gpurun_out/r02ah_bgzf.json')); print(j['cli_hash_sha1']); print(j['cli_all12']); print(j['cli_phases'])"
