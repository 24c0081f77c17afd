mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) 2>&1 | tail -15 > gpurun_out/r15_tests.txt
cat gpurun_out/r15_tests.txt
python bench.py --steps 60 --warmup 3 --no-cpu --hetero > gpurun_out/r15_bench.json 2> gpurun_out/r15_bench.err
tail -3 gpurun_out/r15_bench.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r15_bench.json')); print('base', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['status_bits'], {k:round(v['ms'],0) for k,v in d['kernels'].items()}); print(d.get('heterogeneous'))"
