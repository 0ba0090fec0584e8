set -x
timeout 900 python tools/bench_bgzf.py --fastq --records 3000000 --no-cli > gpurun_out/r02al_bgzf_fastq.json 2> gpurun_out/r02al.err; tail -2 gpurun_out/r02al.err; python -c "
import json; j=json.load(open('gpurun_out/r02al_bgzf_fastq.json')); print({k:j[k] for k in ('records','text_bytes','bgzf_bytes','ratio')}); print(j['device_path']); print(j['slots_4'])"
