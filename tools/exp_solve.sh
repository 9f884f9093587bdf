# solve-only comparison + per-kernel durations of the solve under ncu (profiling only)
for f in 1 0; do SPK_SOLVE_FLOW=$f timeout 300 python tools/run_factor.py --grid 96 --solve 3 2>&1 | grep -E "solve|factor 0" | sed "s/^/flow=$f /"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_p --csv --log-file gpurun_out/ncu_solve_flow.csv python tools/run_factor.py --grid 96 --solve 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/ncu_solve_flow.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size') if 'Grid Size' in hdr else None
agg=collections.OrderedDict()
for r in rows[1:]:
    name=r[ki].split('<')[0].split('(')[0]
    try: v=float(r[vi].replace(',',''))
    except: continue
    print(name, r[gi] if gi is not None else '', v/1e3, 'us')
PY
