#!/bin/bash
# usage: bash tools/gpu_sanitize.sh memcheck|racecheck   (ONE tool per gpurun call: B200_PROFILING.md)
# compute-sanitizer over the tile / fused-strip / fused-hash / long-record / text-path parity tests; the log goes to
# gpurun_out/sanitizer_<tool>.log (copied to profiles/ when clean).
tool=$1
mkdir -p gpurun_out
sel_parity='strip or tile or whitespace or blake3_nthash or goldens or long_record'
sel_text='fixtures or multiline or fastq_chunked or truncated or more_records'
timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_${tool}.log \
    python -m pytest tests/test_parity_gpu.py -x -q -k "$sel_parity" > gpurun_out/sanitizer_${tool}_pytest_parity.log 2>&1
echo "parity rc=$?" | tee -a gpurun_out/sanitizer_${tool}_pytest_parity.log
timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_${tool}_text.log \
    python -m pytest tests/test_textproc_gpu.py -x -q -k "$sel_text" > gpurun_out/sanitizer_${tool}_pytest_text.log 2>&1
echo "text rc=$?" | tee -a gpurun_out/sanitizer_${tool}_pytest_text.log
tail -5 gpurun_out/sanitizer_${tool}.log gpurun_out/sanitizer_${tool}_text.log
tail -3 gpurun_out/sanitizer_${tool}_pytest_parity.log gpurun_out/sanitizer_${tool}_pytest_text.log
