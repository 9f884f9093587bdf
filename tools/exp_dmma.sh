run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dropin --profile > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err; python -c "
import json; d=json.load(open('gpurun_out/x_$name.json')); b=d['breakdown_ms']; print('$name', 'factor_ms %.1f'%(d['factor_s']*1e3), 'dmma128 %.1f dmma64 %.1f asm %.1f'%(b['gemm_dmma_128x64']['ms'], b['gemm_dmma_64x64']['ms'], b['asm']['ms']), 'TF %.2f'%d['roofline']['achieved'])"; }
run base SPK_X=0
run cg SPK_DMMA_CA=0
run nopersist SPK_DMMA_PERSIST=0
run static SPK_DMMA_STATIC=1
run dephase15 SPK_DMMA_DEPHASE=15
run stages4 SPK_DMMA_VARIANT=6
