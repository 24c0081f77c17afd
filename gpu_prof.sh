set -x
mkdir -p gpurun_out
# launch list of 3 frames (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/launches_r1.log 2>&1
# full captures at frame ~40 of the three heavy kernels (one launch each)
for k in k_manifold k_solve k_gjk k_integrate; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 800 -c 1 -o gpurun_out/prof_r1_$k -f python bench.py --steps 42 --warmup 0 --no-extras > gpurun_out/prof_r1_$k.log 2>&1
done
ls -la gpurun_out
