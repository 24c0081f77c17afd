mkdir -p gpurun_out
for spec in k_pos_level:0 k_pos_level:1 k_vel_level:0 k_integrate:10; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s $s -c 1 -o gpurun_out/p9_${k}_$s -f python bench.py --ncu-frame 45 > gpurun_out/p9_${k}_$s.log 2>&1
done
ls -la gpurun_out | grep p9
