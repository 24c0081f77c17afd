set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "trajectory or large or cull" 2>&1 | tail -4
python bench.py --steps 60 --warmup 3 > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_c.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:v['ms'] for k,v in d['kernels'].items()})"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_manifold -s 800 -c 1 -o gpurun_out/prof_r1c_k_manifold -f python bench.py --steps 42 --warmup 0 --no-extras > gpurun_out/prof_r1c.log 2>&1
