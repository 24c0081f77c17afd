# 8-GPU strong-scaling line (4096 worlds in total) of the headline workload over the whole window
T=${1:-scale}; N=${2:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 60 --warmup 3 --scaling strong --no-cpu --e2e-parts 0 \
  > gpurun_out/${T}_n${N}_strong60.json 2> gpurun_out/${T}_n${N}_strong60.err || tail -5 gpurun_out/${T}_n${N}_strong60.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_n${N}_strong60.json')); print('strong60 N=$N', d['scaling'], round(d['value']/1e6,1), 'M bs/s', round(d['ms_per_step'],3), 'ms; e2e', round(d.get('e2e',{}).get('value',0)/1e6,1), 'parity', d.get('parity_checked'), d['parity'].get('worlds'), d['config']['worlds_per_gpu'])"
