mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bigfront_parity.py -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/g_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/g_$name.json 2> gpurun_out/g_$name.err; python -c "
import json; d=json.load(open('gpurun_out/g_$name.json')); print('$name', 'factor_ms %.1f'%(d['factor_s']*1e3), 'solve_ms %.2f'%(d['solve_s']*1e3), 'e2e_ms %.1f'%d['e2e']['ms'], 'TF %.2f'%d['roofline']['achieved'], 'frac %.3f'%d['roofline']['frac'], 'resid %.1e'%d['residual'])"; }
run cfg4 SPK_X=0
run cfg4_inv3 SPK_SOLVE_INV=3
run cfg4_inv1 SPK_SOLVE_INV=1
run cfg2 SPK_BENCH_CONFIG=cfg2
run cfg5 SPK_BENCH_CONFIG=cfg5
