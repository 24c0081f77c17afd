# relaxed-poll dataflow sweeps: GPU tests, headline at K = 20 and over the window with the heterogeneous leg, c2
T=${1:-g4}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) 2>&1 | tail -15 > gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
summ() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', round(d['value']/1e6,2), round(d['ms_per_step'],3), round(d.get('e2e',{}).get('value',0)/1e6,2), {k:round(v['ms'],1) for k,v in d.get('kernels',{}).items()}, d.get('status_bits'), d.get('parity',{}).get('ok'), 'hetero', d.get('heterogeneous',{}).get('ms_per_step'), d.get('heterogeneous',{}).get('kernels_ms'))"; }
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err || tail -5 gpurun_out/${T}_bench20.err
summ gpurun_out/${T}_bench20.json
timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu --hetero > gpurun_out/${T}_bench60.json 2> gpurun_out/${T}_bench60.err || tail -5 gpurun_out/${T}_bench60.err
summ gpurun_out/${T}_bench60.json
timeout 300 python bench.py --workload c2 --warmup 3 --no-cpu > gpurun_out/${T}_c2.json 2> gpurun_out/${T}_c2.err
summ gpurun_out/${T}_c2.json
