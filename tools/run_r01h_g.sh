mkdir -p gpurun_out
M=gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
for bp in 1 0; do
SG_BANKPLAN=$bp SG_STREAMS=1 timeout 600 ncu --metrics $M --clock-control none -k regex:'mesh_v2_kernel|bankplan|graph_kernel' -c 5 --csv --log-file gpurun_out/r01h_g_bp$bp.csv python tools/dp_probe.py --refs 50000 --queries 1184 --reps 1 > gpurun_out/r01h_g_bp$bp.log 2>&1
done
grep -v "^==" gpurun_out/r01h_g_bp1.csv | cut -d, -f5,13- | head -40
grep -v "^==" gpurun_out/r01h_g_bp0.csv | cut -d, -f5,13- | head -40
