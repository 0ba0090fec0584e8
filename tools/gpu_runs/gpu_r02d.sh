set -x
python tools/probe_pipes_ncu.py > gpurun_out/plain_probe.log 2>&1 && \
ncu --metrics sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:probe --csv --log-file gpurun_out/r02d_probe_pipes.csv python tools/probe_pipes_ncu.py > gpurun_out/ncu_probe.log 2>&1
cat gpurun_out/plain_probe.log
timeout 600 python tools/ab_c3.py seqhasher_b200/_variants/lib_u1m1.so seqhasher_b200/_variants/lib_u0m1.so > gpurun_out/r02d_ab.log 2>&1
cat gpurun_out/r02d_ab.log
