mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bigfront_parity.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/e_pytest.log
run() { name=$1; shift; env "$@" SPK_DUMP_LAUNCHES=gpurun_out/e_launches_$name.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --profile > gpurun_out/e_$name.json 2> gpurun_out/e_$name.err; python -c "
import json; d=json.load(open('gpurun_out/e_$name.json')); b=d['breakdown_ms']; print('$name', 'factor_ms %.1f'%(d['factor_s']*1e3), 'solve_ms %.2f'%(d['solve_s']*1e3), 'dmma128 %.1f dmma64 %.1f diag %.1f panel %.1f asm %.1f'%(b['gemm_dmma_128x64']['ms'], b['gemm_dmma_64x64']['ms'], b['diag']['ms'], b['panel']['ms'], b['asm']['ms']), 'TF %.2f'%d['roofline']['achieved'], 'frac %.3f'%d['roofline']['frac'], 'resid %.1e'%d['residual'])"; }
run t64 SPK_X=0
run big SPK_DMMA_BIG=1
run t64_ob12 SPK_OB_STEPS=12
run t64_ob6 SPK_OB_STEPS=6
run t64_old SPK_DMMA_VARIANT64=4
run lu_t64 SPK_BENCH_CONFIG=cfg3
run cfg2_t64 SPK_BENCH_CONFIG=cfg2
run cfg5_t64 SPK_BENCH_CONFIG=cfg5
