set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --steps 60 --warmup 3 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; tail -c 2500 gpurun_out/bench_r1_b.json; tail -5 gpurun_out/bench_r1_b.err
python bench.py --steps 60 --warmup 3 --no-cull --no-extras 2>&1 | tail -c 600
