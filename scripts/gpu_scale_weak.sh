# 8-GPU weak-scaling line of the headline workload only (its `window` block carries the 60-frame figure): `gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_scale_weak.sh TAG 8'`
T=${1:-scale}; N=${2:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu \
  > gpurun_out/${T}_n${N}_weak20.json 2> gpurun_out/${T}_n${N}_weak20.err || tail -5 gpurun_out/${T}_n${N}_weak20.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_n${N}_weak20.json')); print('weak20 N=$N', d['scaling'], round(d['value']/1e6,1), 'M bs/s', round(d['ms_per_step'],3), 'ms; window', round(d['window']['value']/1e6,1), 'e2e', round(d.get('e2e',{}).get('value',0)/1e6,1), 'parity', d.get('parity_checked'), d['parity'].get('worlds'), d['config']['worlds_per_gpu'], d['clocks'])"
