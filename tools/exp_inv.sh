export SPK_SOLVE_GRAPH=0
for iv in 0 3; do
  SPK_SOLVE_INV=$iv timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:"k_p[fb]_step" --launch-skip 900 --launch-count 40 --csv --log-file gpurun_out/ncu_inv$iv.csv python tools/run_factor.py --grid 96 --solve 1 > /dev/null 2>&1
done
python - <<'PY'
import csv, collections
for iv in (0,3):
    rows=[r for r in csv.reader(open(f'gpurun_out/ncu_inv{iv}.csv')) if len(r)>10]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID'); gi=hdr.index('Grid Size')
    d=collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[ii], {'k':r[ki][:22],'g':r[gi]})[r[mi]]=float(r[vi].replace(',',''))
    print('inv',iv)
    for k,v in list(d.items())[:40]:
        print('  ',v['k'],v['g'],'t=%.1f us'%(v['gpu__time_duration.sum']/1e3),'inst=%d'%v['smsp__inst_executed.sum'],'bankconf=%d'%v['l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'])
PY
