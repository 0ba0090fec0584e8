"""Where does the wall time of one CLI invocation go? (startup vs. I/O vs. GPU)"""
import io, os, sys, time
t0 = time.perf_counter()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
t1 = time.perf_counter()
from seqhasher_b200 import binding as B
from seqhasher_b200 import host as H
t2 = time.perf_counter()
B.init(1)
t3 = time.perf_counter()
tp = B.TextProc(64 << 20)
t4 = time.perf_counter()
path = sys.argv[1]
cfg = H.Config(hash_types=["sha1"], headers_only=True, case_sensitive=True, input_file_name=path)
with open(path, "rb") as f, open(os.devnull, "wb") as out:
    H.process_sequences_gpu(f, out, cfg, textproc=tp)
t5 = time.perf_counter()
tp.close()
t6 = time.perf_counter()
print(f"import numpy {t1-t0:.3f}  import pkg {t2-t1:.3f}  shb_init {t3-t2:.3f}  textproc_create {t4-t3:.3f}  process {t5-t4:.3f}  close {t6-t5:.3f}")
