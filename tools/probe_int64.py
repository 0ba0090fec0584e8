"""64-bit integer building blocks of the multiply-based hashes on one B200 (register-only probes, csrc/kernels.cu
probe64_kernel): ops per second where one op = one IMAD.WIDE / IMAD.HI / 64-bit multiply-add / 128-bit fold / rotate / add."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seqhasher_b200 import binding as B  # noqa: E402

B.init(1)
names = {0: "lop3", 3: "imad", 8: "imad_wide_u32", 9: "imad_hi_u32", 10: "mul64_lo_plus_add", 11: "mulfold_128",
         12: "rotl64_31", 13: "add64", 6: "lop3_imad_1to1"}
out = {k: B.probe_int_peak(kind, 200) / 1e12 for kind, k in names.items()}
out["unit"] = "T lane-ops/s (one op = one of the named 32- or 64-bit operations)"
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/INT64_PEAKS.json", "w"), indent=1)
