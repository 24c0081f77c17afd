mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r21_n2.json 2> gpurun_out/r21_n2.err
tail -3 gpurun_out/r21_n2.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r21_n2.json')); print('n2', d['n_gpus'], round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['status_bits'], d.get('aggregate_work'), d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r21_ref_n2.json 2> gpurun_out/r21_ref_n2.err
cut -c1-300 gpurun_out/r21_ref_n2.json; tail -2 gpurun_out/r21_ref_n2.err
