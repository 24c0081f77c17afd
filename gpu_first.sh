set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 2 --worlds 512 2>&1 | tail -3
python bench.py --steps 60 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; tail -c 3000 gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
