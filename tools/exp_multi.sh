# usage: exp_multi.sh N "k1 k2 ..."   (SPK_TOP_SPLITS values; "d" = the cost model's own choice)
N=$1
for k in $2; do
  if [ "$k" != "d" ]; then export SPK_TOP_SPLITS=$k; else unset SPK_TOP_SPLITS; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m_n${N}_k$k.json 2> gpurun_out/m_n${N}_k$k.err
  python -c "
import json; d=json.load(open('gpurun_out/m_n${N}_k$k.json')); print('N=$N splits=$k', 'TF %.1f'%(d['value']/1e3), 'factor_ms %.1f'%(d['factor_s']*1e3), 'phases', [round(x,1) for x in d['phase_ms']], 'solve_ms %.1f'%(d['solve_s']*1e3), d['parallelism'][:24], 'GiB %.1f'%(d['device_bytes']/2**30), 'resid %.1e'%d['residual'], d.get('rank_phase_ms'))"
done
