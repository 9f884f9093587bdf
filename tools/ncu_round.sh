# profiles for the round: (1) launch list of one factorisation + solve, (2) --set full of the dominant kernels,
# exported on the box to small CSVs (raw page) so that gpurun_out stays small
export SPK_LOOKAHEAD=0 SPK_SOLVE_GRAPH=0
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_full.csv python tools/run_factor.py --grid 96 --reps 1 --solve 1 > gpurun_out/r01_launches.log 2>&1
ncu --metrics $M --clock-control none --kernel-name regex:k_gemm_dmma --launch-skip 250 --launch-count 12 --csv --log-file gpurun_out/r01_ncu_dmma.csv python tools/run_factor.py --grid 96 --reps 1 > /dev/null 2>&1
ncu --metrics $M --clock-control none --kernel-name regex:k_assemble --launch-skip 22 --launch-count 8 --csv --log-file gpurun_out/r01_ncu_asm.csv python tools/run_factor.py --grid 96 --reps 1 > /dev/null 2>&1
ncu --metrics $M --clock-control none --kernel-name regex:"k_pf_step|k_pb_step|k_pf_front|k_pb_front" --launch-count 60 --csv --log-file gpurun_out/r01_ncu_solve.csv python tools/run_factor.py --grid 96 --reps 1 --solve 1 > /dev/null 2>&1
ncu --metrics $M --clock-control none --kernel-name regex:"k_panel_reg|k_diag_ldlt_row|k_chunks|k_scatter" --launch-count 30 --csv --log-file gpurun_out/r01_ncu_step.csv python tools/run_factor.py --grid 96 --reps 1 > /dev/null 2>&1
gzip -f gpurun_out/r01_launches_full.csv
ls -la gpurun_out | tail -8
