set -x
(timeout 900 python -m pytest tests/test_bgzf_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/r02z_pytest.log
tail -2 gpurun_out/r02z_pytest.log
timeout 900 python tools/bench_bgzf.py --no-cli --records 1000000 > gpurun_out/r02z_bgzf.json 2> gpurun_out/r02z_bgzf.err; tail -2 gpurun_out/r02z_bgzf.err; python -c "
import json; j=json.load(open('gpurun_out/r02z_bgzf.json')); print(j['device_path']); print(j['slots_2']); print(j['slots_4'])"
