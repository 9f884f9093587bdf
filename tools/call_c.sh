# call C (1 GPU): parity of the split trailing update, bench on/off, per-level dumps for several outer-block sizes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bigfront_parity.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c_pytest.log
run() { name=$1; shift; env "$@" SPK_DUMP_LAUNCHES=gpurun_out/c_launches_$name.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --profile > gpurun_out/c_$name.json 2> gpurun_out/c_$name.err; python -c "
import json; d=json.load(open('gpurun_out/c_$name.json')); b=d['breakdown_ms']; print('$name', 'factor_ms %.1f'%(d['factor_s']*1e3), 'solve_ms %.2f'%(d['solve_s']*1e3), 'dmma128 %.1f dmma64 %.1f diag %.1f panel %.1f asm %.1f'%(b['gemm_dmma_128x64']['ms'], b['gemm_dmma_64x64']['ms'], b['diag']['ms'], b['panel']['ms'], b['asm']['ms']), 'TF %.2f'%d['roofline']['achieved'], 'resid %.1e'%d['residual'])"; }
run split SPK_X=0
run nosplit SPK_SPLIT_REST=0
run split_ob12 SPK_OB_STEPS=12
run split_ob4 SPK_OB_STEPS=4
run split_ob6 SPK_OB_STEPS=6
run split_ob16 SPK_OB_STEPS=16
run lu_split SPK_BENCH_CONFIG=cfg3
run lu_nosplit SPK_BENCH_CONFIG=cfg3 SPK_SPLIT_REST=0
SPK_TRACE=gpurun_out/c_trace_cfg4.csv timeout 200 python tools/run_factor.py --grid 96 --reps 2 > gpurun_out/c_trace.log 2>&1
