# one ncu --set full capture of the seven substep kernels at frame 40 of the headline workload (source-level counters included)
T=${1:-ncu}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_(integrate|cull|gjk|epa|manifold|solve_pos|solve_vel)' -c 7 -o gpurun_out/${T}_substep -f python bench.py --ncu-frame 40 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
ncu -i gpurun_out/${T}_substep.ncu-rep --page raw --csv > gpurun_out/${T}_substep.raw.csv 2>/dev/null
ls -la gpurun_out/${T}_substep.ncu-rep
