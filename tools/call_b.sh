# call B (1 GPU): cuBLAS on the update shapes, real two-stream trace of cfg4, outer-block-size / look-ahead sweep
mkdir -p gpurun_out
timeout 120 python tools/cublas_shapes.py > gpurun_out/b_cublas_shapes.txt 2>&1
SPK_TRACE=gpurun_out/b_trace_cfg4.csv timeout 200 python tools/run_factor.py --grid 96 --reps 2 > gpurun_out/b_trace.log 2>&1
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --profile > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err; python -c "
import json; d=json.load(open('gpurun_out/b_$name.json')); b=d['breakdown_ms']; print('$name', 'factor_ms %.1f'%(d['factor_s']*1e3), 'solve_ms %.2f'%(d['solve_s']*1e3), 'dmma128 %.1f dmma64 %.1f diag %.1f panel %.1f asm %.1f'%(b['gemm_dmma_128x64']['ms'], b['gemm_dmma_64x64']['ms'], b['diag']['ms'], b['panel']['ms'], b['asm']['ms']), 'TF %.2f'%d['roofline']['achieved'], 'resid %.1e'%d['residual'])"; }
run ob4 SPK_OB_STEPS=4
run ob6 SPK_OB_STEPS=6
run ob12 SPK_OB_STEPS=12
run ob16 SPK_OB_STEPS=16
run nola SPK_LOOKAHEAD=0
run inv3 SPK_SOLVE_INV=3
