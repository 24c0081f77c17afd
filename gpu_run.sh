mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) 2>&1 | tail -8 > gpurun_out/r26_tests.txt
cat gpurun_out/r26_tests.txt
python bench.py --steps 60 --warmup 3 --no-cpu --hetero > gpurun_out/r26_bench.json 2> gpurun_out/r26_bench.err
tail -3 gpurun_out/r26_bench.err
python -c "
import json,sys
d=json.load(open('gpurun_out/r26_bench.json')); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step'],2), d.get('e2e_single_batch',{}).get('ms_per_step'), d.get('e2e_two_half_batches'), d['heterogeneous']['ms_per_step'])"
