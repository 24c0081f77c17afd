# multi-GPU evidence: `gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_scale.sh TAG N'` (one process per GPU under torchrun, NCCL)
T=${1:-scale}; N=${2:-8}
mkdir -p gpurun_out
run() {  # name, bench args...
  NAME=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" \
    > gpurun_out/${T}_n${N}_$NAME.json 2> gpurun_out/${T}_n${N}_$NAME.err || tail -5 gpurun_out/${T}_n${N}_$NAME.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_n${N}_$NAME.json')); print('$NAME N=$N', d['scaling'], round(d['value']/1e6,1), 'M bs/s', round(d['ms_per_step'],3), 'ms; window', round(d['window']['value']/1e6,1), 'e2e', round(d.get('e2e',{}).get('value',0)/1e6,1), 'parity', d.get('parity_checked'), d['parity'].get('worlds'), d['config']['worlds_per_gpu'])"
}
run weak20 --steps 20 --warmup 3 --no-cpu
run strong60 --steps 60 --warmup 3 --scaling strong --no-cpu --e2e-parts 0
run c5 --workload c5 --warmup 3 --no-cpu --e2e-parts 0
run c2 --workload c2 --warmup 3 --no-cpu --e2e-parts 0
