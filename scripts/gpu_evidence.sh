# end-of-round evidence: bench lines of both arms, launch list, ncu captures (summarise with profiles/summarise_ncu.py)
T=${1:-final}
mkdir -p gpurun_out
python bench.py --steps 60 --warmup 3 --hetero > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 60 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
cut -c1-200 gpurun_out/${T}_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 60 --warmup 3 --no-extras > gpurun_out/${T}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_(integrate|cull|transform|gjk|epa|manifold|solve_pos|solve_vel)' -c 8 -o gpurun_out/${T}_substep -f python bench.py --ncu-frame 40 > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_substep.ncu-rep --page raw --csv > gpurun_out/${T}_substep.raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'k_(broad_cells|broad_scan|broad_write|islands|schedule)' -c 5 -o gpurun_out/${T}_prologue -f python bench.py --ncu-frame 40 > gpurun_out/${T}_ncu2.log 2>&1
ncu -i gpurun_out/${T}_prologue.ncu-rep --page raw --csv > gpurun_out/${T}_prologue.raw.csv 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json')); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['roofline']['kernel'], round(d['roofline']['frac'],3), round(d['fp64']['whole_step_frac'],3), round(d['heterogeneous']['ms_per_step'],1))"
ls gpurun_out | grep ${T}_
