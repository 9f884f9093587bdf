export SPK_DIAG_TG=1
ncu --set full --import-source on --clock-control none --kernel-name regex:k_diag_ldlt_row --launch-skip 140 --launch-count 1 -o gpurun_out/diag_row -f python tools/run_factor.py --grid 48 --reps 1 > gpurun_out/ncu_diag.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:k_panel_reg --launch-skip 140 --launch-count 1 -o gpurun_out/panel_reg -f python tools/run_factor.py --grid 48 --reps 1 > gpurun_out/ncu_panel.log 2>&1
tail -3 gpurun_out/ncu_diag.log
tail -3 gpurun_out/ncu_panel.log
