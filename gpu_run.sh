mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) 2>&1 | tail -15 > gpurun_out/r22_tests.txt
cat gpurun_out/r22_tests.txt
python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/r22_bench.json 2> gpurun_out/r22_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r22_bench20.json 2> gpurun_out/r22_bench20.err
wc -l gpurun_out/r22_bench.json gpurun_out/r22_bench20.json
python -c "
import json,sys
for f in ('r22_bench','r22_bench20'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['status_bits'], d['config']['workload'], {k:round(v['ms'],0) for k,v in d['kernels'].items()})"
H=raw-physics_b200/rp_headless
$H --scene stack --worlds 4096 --frames 600
$H --scene brick_wall --rows 32 --cols 32 --worlds 1 --frames 60
$H --scene brick_wall --rows 32 --cols 32 --worlds 64 --frames 60
$H --scene levers --worlds 16384 --frames 120
$H --scene w256 --worlds 4096 --frames 60
