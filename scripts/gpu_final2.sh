# evidence of the final build, trimmed to what changed since scripts/gpu_final.sh ran: GPU tests, bench lines, one ncu capture of the substep
T=${1:-final2}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 > gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err; tail -2 gpurun_out/${T}_bench20.err
python bench.py --steps 60 --warmup 3 --hetero --no-cpu > gpurun_out/${T}_bench60.json 2> gpurun_out/${T}_bench60.err; tail -2 gpurun_out/${T}_bench60.err
for WL in c2 c3 c5 pile; do
  timeout 600 python bench.py --workload $WL --warmup 3 --no-cpu > gpurun_out/${T}_$WL.json 2> gpurun_out/${T}_$WL.err || tail -3 gpurun_out/${T}_$WL.err
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_(integrate|cull|gjk|epa|manifold|solve_pos|solve_vel)' -c 7 -o gpurun_out/${T}_substep -f python bench.py --ncu-frame 40 > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_substep.ncu-rep --page raw --csv > gpurun_out/${T}_substep.raw.csv 2>/dev/null
python -c "
import json
for k in ['bench20','bench60','c2','c3','c5','pile']:
    d=json.load(open('gpurun_out/${T}_%s.json' % k)); print(k, round(d['value']/1e6,2), round(d['ms_per_step'],3), round(d['e2e']['value']/1e6,2), d['parity_checked'], d['roofline']['kernel'], round(d['fp64']['whole_step_frac'],3), d.get('heterogeneous',{}).get('ms_per_step'))"
