# the other BASELINE configurations through bench.py: `gpurun --timeout 900 -- 'bash scripts/gpu_workloads.sh TAG'`
T=${1:-wl}
mkdir -p gpurun_out
for WL in ${2:-c2 c3 c5 pile}; do
  timeout 400 python bench.py --workload $WL --warmup 3 > gpurun_out/${T}_$WL.json 2> gpurun_out/${T}_$WL.err || tail -5 gpurun_out/${T}_$WL.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_$WL.json')); print('$WL', round(d['value']/1e6,2), 'M bs/s', round(d['ms_per_step'],3), 'ms; e2e', round(d.get('e2e',{}).get('value',0)/1e6,2), 'parity', d.get('parity_checked'), d.get('parity',{}).get('max_abs_pose_diff_vs_reference'), {k:round(v['ms'],1) for k,v in d.get('kernels',{}).items()}, 'cpu', round(d.get('cpu_baseline',{}).get('value',0)/1e6,3), d.get('reference_order',{}).get('ms_per_step'))"
done
