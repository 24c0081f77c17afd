# what a development round trip runs on the GPU box: `gpurun --timeout 900 -- 'bash scripts/gpu_check.sh'`
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -8 > gpurun_out/check_tests.txt
cat gpurun_out/check_tests.txt
timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err
tail -3 gpurun_out/check_bench.err
python -c "
import json,sys
d=json.load(open('gpurun_out/check_bench.json')); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],0) for k,v in d['kernels'].items()})"
