mkdir -p gpurun_out
( for k in 64 200 400; do timeout 120 ./tools/ubench_dmma 13000 $k 0 0 3 64; done; timeout 120 ./tools/ubench_dmma 4000 400 0 0 3 64; timeout 120 ./tools/ubench_dmma 13000 456 0 0 3 512;  timeout 120 ./tools/ubench_dmma 9000 456 1 0 3 ) > gpurun_out/d_ubench4.txt 2>&1
cat gpurun_out/d_ubench4.txt
