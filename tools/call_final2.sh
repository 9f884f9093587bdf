# final 2-GPU validation: multi-GPU parity tests + the N=2 bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/z_pytest_gpu2_multi.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/z_pytest_gpu2_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/z_bench_n2.json 2> gpurun_out/z_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/z_bench_n2.json"))
print("N=2 TF %.2f" % (d["value"]/1e3), "factor_ms %.1f" % (d["factor_s"]*1e3), "solve_ms %.2f" % (d["solve_s"]*1e3), "e2e %.2f" % (d["e2e"]["value"]/1e3), "GiB %.1f" % (d["device_bytes"]/2**30), "resid %.1e" % d["residual"], d.get("phase_ms"), d.get("rank_phase_ms"), d["parallelism"])
PY
