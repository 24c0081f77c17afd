# one multi-GPU bench line the way the driver launches it: `gpurun --gpus N --timeout 600 -- 'bash scripts/gpu_scale.sh N'`
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 60 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
tail -2 gpurun_out/scale_n$N.err
wc -l gpurun_out/scale_n$N.json
python -c "
import json,sys; d=json.load(open('gpurun_out/scale_n$N.json')); print('n', d['n_gpus'], round(d['value']/1e9,3), 'e9', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value']/1e9,3), d['status_bits'], d['clocks'])"
