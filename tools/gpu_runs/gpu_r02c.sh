set -x
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r02c_probes.log 2>&1
import sys, json
sys.path.insert(0, ".")
from seqhasher_b200 import binding as B
B.init(1)
names = {0: "lop3", 3: "imad", 6: "lop3:imad 1:1", 34: "lop3:viadd 1:1", 33: "imad:iadd3 1:1 (add-imm)", 35: "imad:iadd3 1:1 (add reg)",
         36: "lop3:add-reg 1:1", 37: "lop3:imad 2:1", 38: "lop3:imad 1:2", 39: "lop3:viadd:imad:iadd3 1:1:1:1"}
print(json.dumps({v: B.probe_int_peak(k, 200) / 1e12 for k, v in names.items()}, indent=1))
PY
cat gpurun_out/r02c_probes.log
(timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_soak_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/r02c_pytest.log
cat gpurun_out/r02c_pytest.log
timeout 900 python tools/ab_c3.py seqhasher_b200/libseqhash_b200.so seqhasher_b200/_variants/lib_u1m0.so seqhasher_b200/_variants/lib_u0m0.so seqhasher_b200/_variants/lib_u2m0.so seqhasher_b200/_variants/lib_u2m1.so seqhasher_b200/_variants/lib_u1m1.so > gpurun_out/r02c_ab.log 2>&1
cat gpurun_out/r02c_ab.log
