mkdir -p gpurun_out
UB_ONLY="64x64 tk8 s4" UB_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dmma --launch-skip 1 --launch-count 1 -o gpurun_out/f_dmma64 ./tools/ubench_dmma 9000 456 1 0 3 > gpurun_out/f_ncu.log 2>&1
ncu -i gpurun_out/f_dmma64.ncu-rep --page raw --csv > gpurun_out/f_dmma64_raw.csv 2>/dev/null
ncu -i gpurun_out/f_dmma64.ncu-rep --page source --csv > gpurun_out/f_dmma64_source.csv 2>/dev/null
ncu -i gpurun_out/f_dmma64.ncu-rep --page details > gpurun_out/f_dmma64_details.txt 2>/dev/null
rm -f gpurun_out/f_dmma64.ncu-rep
tail -5 gpurun_out/f_ncu.log; ls -la gpurun_out/f_*
