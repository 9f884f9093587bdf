mkdir -p gpurun_out
( timeout 120 ./tools/ubench_dmma 9000 456 1; timeout 120 ./tools/ubench_dmma 9000 456 1 0 3; timeout 120 ./tools/ubench_dmma 4000 456 1; timeout 120 ./tools/ubench_dmma 9000 200 1;  timeout 120 ./tools/ubench_dmma 2000 456 1) > gpurun_out/d_ubench3.txt 2>&1
cat gpurun_out/d_ubench3.txt
