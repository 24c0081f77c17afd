mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4
python bench.py --steps 60 --warmup 3 > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_f.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:v['ms'] for k,v in d['kernels'].items()})"
tail -3 gpurun_out/bench_r1_f.err
