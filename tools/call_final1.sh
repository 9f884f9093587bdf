# final single-GPU validation of the round: full GPU test-suite (multi-GPU tests skip on one GPU), smoke, bench lines of all five configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/z_pytest_1gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/z_pytest_1gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/z_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --profile > gpurun_out/z_bench_cfg4.json 2> gpurun_out/z_bench_cfg4.err; echo "cfg4 rc=$?"
for c in cfg1 cfg2 cfg3 cfg5; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/z_bench_$c.json 2> gpurun_out/z_bench_$c.err; echo "$c rc=$?"; done
python - <<'PY'
import json
for c in ("cfg1","cfg2","cfg3","cfg4","cfg5"):
    try:
        d = json.load(open(f"gpurun_out/z_bench_{c}.json"))
        print(c, "TF %.2f" % (d["value"]/1e3), "factor_ms %.2f" % (d["factor_s"]*1e3), "solve_ms %.2f" % (d["solve_s"]*1e3), "e2e_ms %.1f (%.2f TF)" % (d["e2e"]["ms"], d["e2e"]["value"]/1e3),
              "dropin_ms %.0f" % d["e2e_dropin"]["ms"], "frac %.3f" % (d["roofline"]["frac"] or 0), "GiB %.1f" % (d["device_bytes"]/2**30), "resid %.1e" % d["residual"], "cpu %.1f GF" % d["cpu_baseline"]["value"], d["clocks"])
    except Exception as e:
        print(c, "FAILED", e)
PY
