mkdir -p gpurun_out
python bench.py --steps 60 --warmup 3 --hetero > gpurun_out/r34_bench.json 2> gpurun_out/r34_bench.err
tail -3 gpurun_out/r34_bench.err
python -c "
import json,sys
d=json.load(open('gpurun_out/r34_bench.json')); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['roofline'], d['fp64'], d['cpu_baseline'])"
python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/r34_ref.json 2> gpurun_out/r34_ref.err
cut -c1-330 gpurun_out/r34_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r34_launches.csv python bench.py --steps 60 --warmup 3 --no-extras > gpurun_out/r34_launches.log 2>&1
wc -l gpurun_out/r34_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_(integrate|cull|gjk|epa|manifold|solve_pos|solve_vel)' -c 7 -o gpurun_out/r34_substep -f python bench.py --ncu-frame 40 > gpurun_out/r34_ncu.log 2>&1
tail -2 gpurun_out/r34_ncu.log
ncu -i gpurun_out/r34_substep.ncu-rep --page raw --csv > gpurun_out/r34_substep.raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'k_(broad_cells|broad_scan|broad_write|islands|schedule)' -c 5 -o gpurun_out/r34_prologue -f python bench.py --ncu-frame 40 > gpurun_out/r34_ncu2.log 2>&1
ncu -i gpurun_out/r34_prologue.ncu-rep --page raw --csv > gpurun_out/r34_prologue.raw.csv 2>/dev/null
ls -la gpurun_out | grep r34
