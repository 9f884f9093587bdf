# Round-2 (second half) profiles after the switch to 64 x 64 DMMA tiles (one GPU, under gpurun; numbers printed under ncu are never bench values):
#  (1) launch list of one 96^3 factorisation + one solve (gpu__time_duration per launch)
#  (2) time, DRAM bytes, tensor-pipe activity, occupancy, L2 hit rate for ALL k_gemm_dmma launches (-> dmma_traffic.json for bench.py's roofline.traffic)
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct"
export SPK_LOOKAHEAD=0 SPK_SOLVE_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_full.csv python tools/run_factor.py --grid 96 --reps 1 --solve 1 > gpurun_out/r02b_launches.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --kernel-name regex:k_gemm_dmma --csv --log-file gpurun_out/r02b_ncu_dmma_all.csv python tools/run_factor.py --grid 96 --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections, json
def rows(path):
    out=[]; hdr=None
    for r in csv.reader(open(path, errors='ignore')):
        if hdr is None:
            if r and r[0]=='ID': hdr=r
            continue
        if len(r)==len(hdr): out.append(dict(zip(hdr,r)))
    return out
R=rows('gpurun_out/r02b_launches_full.csv')
agg=collections.defaultdict(lambda:[0,0.0])
for r in R:
    if r['Metric Name']!='gpu__time_duration.sum': continue
    name=r['Kernel Name'].split('<')[0].split('(')[0].replace('void ','').replace('spk::','')
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    ms = v/1e6 if u in ('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
    agg[name][0]+=1; agg[name][1]+=ms
tot=sum(v[1] for v in agg.values())
with open('gpurun_out/r02b_launches_summary.csv','w') as f:
    f.write('kernel,launches,total_ms,share_pct\n')
    for k,v in sorted(agg.items(), key=lambda x:-x[1][1]): f.write(f'{k},{v[0]},{v[1]:.3f},{100*v[1]/tot:.2f}\n')
print(open('gpurun_out/r02b_launches_summary.csv').read())
D=rows('gpurun_out/r02b_ncu_dmma_all.csv')
per=collections.defaultdict(dict)
for r in D: per[r['ID']][r['Metric Name']]=float(r['Metric Value'].replace(',',''))
units={r['Metric Name']:r['Metric Unit'] for r in D}
n=len(per); byt=sum(v.get('dram__bytes_read.sum',0)+v.get('dram__bytes_write.sum',0) for v in per.values())
ub=units.get('dram__bytes_read.sum','byte'); scale={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(ub,1)
t=sum(v['gpu__time_duration.sum'] for v in per.values()); ut=units['gpu__time_duration.sum']; tms=t/1e6 if ut.startswith('n') else t/1e3
tw=sum(v['gpu__time_duration.sum']*v['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'] for v in per.values())/t
print('dmma launches',n,'total ms',tms,'dram bytes',byt*scale,'per launch',byt*scale/n,'time-weighted tensor pipe active %',tw, 'units', ub, ut)
json.dump({'cfg4':{'bytes_per_launch':byt*scale/n,'launches':n,'total_bytes':byt*scale,'total_ms_under_ncu':tms,'tensor_pipe_active_pct_time_weighted':tw,
  'source':'profiles/r02b_ncu_dmma_all_96cubed.csv.gz (ncu dram__bytes_read.sum + dram__bytes_write.sum over all k_gemm_dmma launches of one 96^3 factorisation, tools/ncu_round2b.sh)'}}, open('gpurun_out/dmma_traffic.json','w'), indent=1)
PY
gzip -f gpurun_out/r02b_launches_full.csv gpurun_out/r02b_ncu_dmma_all.csv
unset SPK_LOOKAHEAD SPK_SOLVE_GRAPH
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/h_$name.json 2> gpurun_out/h_$name.err; python -c "
import json; d=json.load(open('gpurun_out/h_$name.json')); print('$name', 'factor_ms %.1f'%(d['factor_s']*1e3), 'solve_ms %.2f'%(d['solve_s']*1e3), 'e2e_ms %.1f'%d['e2e']['ms'], 'TF %.2f'%d['roofline']['achieved'], 'resid %.1e'%d['residual'])"; }
run pipes2 SPK_PIPES=2
run pipes3 SPK_PIPES=3
run reserve SPK_DMMA_PERSIST=1 SPK_GEMM_RESERVE=32
run pdlf SPK_PDL_FACTOR=1
