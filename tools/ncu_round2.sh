# Round-2 profiles (one GPU, under gpurun; numbers printed under ncu are never bench values):
#  (1) launch list of one 96^3 factorisation + one solve (gpu__time_duration per launch)
#  (2) selected metrics (time, DRAM bytes, tensor pipe, occupancy, L2 hit rate) for the DMMA update kernel: the first 40 launches of the
#      top two levels + a sample of mid-tree launches; per-launch DRAM traffic of ALL DMMA launches for bench.py's roofline.traffic
#  (3) one --set full capture of three large DMMA launches (source-level stalls), exported to CSV on the box
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active"
export SPK_LOOKAHEAD=0 SPK_SOLVE_GRAPH=0
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_full.csv python tools/run_factor.py --grid 96 --reps 1 --solve 1 > gpurun_out/r02_launches.log 2>&1
ncu --metrics $M --clock-control none --kernel-name regex:k_gemm_dmma --csv --log-file gpurun_out/r02_ncu_dmma_all.csv python tools/run_factor.py --grid 96 --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_gemm_dmma --launch-skip 560 --launch-count 3 -o gpurun_out/r02_dmma_full python tools/run_factor.py --grid 96 --reps 1 > /dev/null 2>&1
ncu -i gpurun_out/r02_dmma_full.ncu-rep --page raw --csv > gpurun_out/r02_dmma_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_dmma_full.ncu-rep --page source --csv > gpurun_out/r02_dmma_full_source.csv 2>/dev/null
rm -f gpurun_out/r02_dmma_full.ncu-rep
gzip -f gpurun_out/r02_launches_full.csv gpurun_out/r02_dmma_full_source.csv
ls -la gpurun_out | grep r02_ | tail -8
