set -x
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
for spec in k_manifold:800 k_integrate:800 k_pos_level:3000 k_gjk:800 k_vel_level:3000; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/p2_$k -f python bench.py --steps 42 --warmup 0 --no-extras > gpurun_out/p2_$k.log 2>&1
  ncu -i gpurun_out/p2_$k.ncu-rep --page source --csv > gpurun_out/p2_$k.source.csv 2>/dev/null
  ncu -i gpurun_out/p2_$k.ncu-rep --page raw --csv > gpurun_out/p2_$k.raw.csv 2>/dev/null
done
ls -la gpurun_out
du -sm gpurun_out
