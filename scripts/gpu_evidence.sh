# end-of-round evidence: bench lines of both arms at the driver's setting and over the whole window, the other workloads,
# launch list, ncu captures (summarise with profiles/summarise_ncu.py / profiles/make_summary.py)
T=${1:-final}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err; tail -2 gpurun_out/${T}_bench20.err
python bench.py --steps 60 --warmup 3 --hetero > gpurun_out/${T}_bench60.json 2> gpurun_out/${T}_bench60.err; tail -2 gpurun_out/${T}_bench60.err
python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/${T}_ref20.json 2> gpurun_out/${T}_ref20.err
for WL in c2 c3 c5 pile; do
  timeout 600 python bench.py --workload $WL --warmup 3 > gpurun_out/${T}_$WL.json 2> gpurun_out/${T}_$WL.err || tail -3 gpurun_out/${T}_$WL.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/${T}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_(integrate|cull|gjk|epa|manifold|solve_pos|solve_vel)' -c 7 -o gpurun_out/${T}_substep -f python bench.py --ncu-frame 40 > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_substep.ncu-rep --page raw --csv > gpurun_out/${T}_substep.raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'k_(broad_cells|broad_scan|broad_write|islands|schedule)' -c 5 -o gpurun_out/${T}_prologue -f python bench.py --ncu-frame 40 > gpurun_out/${T}_ncu2.log 2>&1
ncu -i gpurun_out/${T}_prologue.ncu-rep --page raw --csv > gpurun_out/${T}_prologue.raw.csv 2>/dev/null
python -c "
import json
for k in ['bench20','bench60','c2','c3','c5','pile']:
    d=json.load(open('gpurun_out/${T}_%s.json' % k)); print(k, round(d['value']/1e6,1), round(d['ms_per_step'],3), round(d['e2e']['value']/1e6,1), d['parity_checked'], d['roofline']['kernel'], round(d['roofline']['frac'],3), round(d['fp64']['whole_step_frac'],3), round(d['cpu_baseline']['value']/1e3,1))
d=json.load(open('gpurun_out/${T}_ref20.json')); print('ref', round(d['value']/1e6,2))"
ls gpurun_out | grep ${T}_ | wc -l
