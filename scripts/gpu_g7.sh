# body-level dataflow sweeps: GPU tests, headline K = 20, the window with the heterogeneous leg, c2, c5 in both forms
T=${1:-g7}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 > gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
summ() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', round(d['value']/1e6,2), round(d['ms_per_step'],3), {k:round(v['ms'],1) for k,v in d.get('kernels',{}).items() if k.startswith('solve') or k in ('manifold','integrate')}, d.get('status_bits'), d.get('parity',{}).get('ok'), 'hetero', d.get('heterogeneous',{}).get('ms_per_step'), {k:v for k,v in (d.get('heterogeneous',{}).get('kernels_ms') or {}).items() if k.startswith('solve')})"; }
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-parts 0 > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err || tail -5 gpurun_out/${T}_bench20.err
summ gpurun_out/${T}_bench20.json
timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu --hetero --e2e-parts 0 > gpurun_out/${T}_bench60.json 2> gpurun_out/${T}_bench60.err || tail -5 gpurun_out/${T}_bench60.err
summ gpurun_out/${T}_bench60.json
timeout 300 python bench.py --workload c2 --warmup 3 --no-cpu --e2e-parts 0 > gpurun_out/${T}_c2.json 2> gpurun_out/${T}_c2.err
summ gpurun_out/${T}_c2.json
RP_FLOW=2 timeout 300 python bench.py --workload c5 --warmup 3 --no-cpu --e2e-parts 0 > gpurun_out/${T}_c5flow.json 2> gpurun_out/${T}_c5flow.err
summ gpurun_out/${T}_c5flow.json
RP_FLOW=2 timeout 300 python bench.py --workload c3 --warmup 3 --no-cpu --e2e-parts 0 > gpurun_out/${T}_c3flow.json 2> gpurun_out/${T}_c3flow.err
summ gpurun_out/${T}_c3flow.json
