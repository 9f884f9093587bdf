mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bigfront_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/n_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --profile > gpurun_out/n_$name.json 2> gpurun_out/n_$name.err; python -c "
import json; d=json.load(open('gpurun_out/n_$name.json')); b=d['breakdown_ms']; print('$name', 'factor_ms %.1f'%(d['factor_s']*1e3), 'solve_ms %.2f'%(d['solve_s']*1e3), 'dmma64 %.1f (%d launches) diag %.1f panel %.1f'%(b['gemm_dmma_64x64']['ms'], b['gemm_dmma_64x64']['launches'], b['diag']['ms'], b['panel']['ms']), 'TF %.2f'%d['roofline']['achieved'], 'resid %.1e'%d['residual'])"; }
run narrow SPK_X=0
run nonarrow SPK_DMMA_NARROW=0
run narrow592 SPK_DMMA_NARROW=592
