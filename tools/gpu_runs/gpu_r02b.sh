set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02b_pytest.log
timeout 600 python bench.py --only c3 > gpurun_out/r02b_bench_c3.json 2> gpurun_out/r02b_bench.err
tail -3 gpurun_out/r02b_bench.err
python - <<'PY' > gpurun_out/r02b_cli_timing.log 2>&1
import numpy as np, subprocess, time, os
rng = np.random.default_rng(1)
n = 400_000
lens = rng.integers(30, 3001, size=n)
seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(lens.sum())).tobytes()
parts, pos = [], 0
for i, L in enumerate(lens):
    parts.append(b">seq_%d\n" % i); parts.append(seq[pos:pos+L]); parts.append(b"\n"); pos += int(L)
open("/tmp/t.fa", "wb").write(b"".join(parts))
print("file bytes", os.path.getsize("/tmp/t.fa"))
env = dict(os.environ, SHB_TIMING="1")
for args in (["--headersonly"], ["--headersonly", "--hash", "sha1,sha3,md5,xxhash,xxh3,cityhash,murmur3,nthash,blake3,k12,rapidhash,rapidhash32"]):
    for rep in range(3):
        t0 = time.perf_counter()
        p = subprocess.run(["seqhasher_b200/bin/seqhasher-b200"] + args + ["/tmp/t.fa", "/dev/null"], capture_output=True, env=env)
        print(args[-1][:20], rep, "wall %.3f" % (time.perf_counter() - t0), p.returncode, p.stderr.decode().replace("\n", " | "))
PY
timeout 300 python tools/bench_configs.py --only C3 --scale 0.2 --steps 2 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hash_tile_kernel -s 3 -c 1 -o gpurun_out/r02b_c3_tile python tools/bench_configs.py --only C3 --scale 0.2 --steps 2 > gpurun_out/ncu_c3.log 2>&1
cat gpurun_out/plain_c3.log | tail -2
