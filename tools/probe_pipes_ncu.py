"""Which pipe takes which integer instruction?  Run single probes for ncu:
ncu --metrics sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed python tools/probe_pipes_ncu.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seqhasher_b200 import binding as B  # noqa: E402

B.init(1)
for kind in (0, 3, 34, 36, 8, 9, 10, 12, 13):
    print(kind, B.probe_int_peak(kind, 20) / 1e12, flush=True)
