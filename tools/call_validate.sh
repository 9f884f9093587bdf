# Round-2 validation call (2 GPUs): full GPU test-suite, 1-GPU bench (cfg4 with per-launch dump, cfg3), 2-GPU bench.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/c1_pytest.log
SPK_DUMP_LAUNCHES=gpurun_out/c1_launches_cfg4.csv timeout 600 python bench.py --steps 3 --warmup 3 --profile --no-cpu-baseline > gpurun_out/c1_bench_cfg4.json 2> gpurun_out/c1_bench_cfg4.err; echo "bench cfg4 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_n2.json 2> gpurun_out/c1_bench_n2.err; echo "bench n2 rc=$?"
SPK_DUMP_LAUNCHES=gpurun_out/c1_launches_cfg3.csv timeout 400 python bench.py --config cfg3 --steps 3 --warmup 3 --profile --no-cpu-baseline --no-dropin > gpurun_out/c1_bench_cfg3.json 2> gpurun_out/c1_bench_cfg3.err; echo "bench cfg3 rc=$?"
python - <<'PY'
import json
for f in ("c1_bench_cfg4","c1_bench_n2","c1_bench_cfg3"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "TF %.2f" % (d["value"]/1e3), "factor_ms %.1f" % (d["factor_s"]*1e3), "solve_ms %.2f" % (d["solve_s"]*1e3), "e2e %.2f" % (d["e2e"]["value"]/1e3),
              "frac %.3f" % (d["roofline"]["frac"] or 0), "GiB %.1f" % (d["device_bytes"]/2**30), "resid %.1e" % d["residual"], d.get("phase_ms"), d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
