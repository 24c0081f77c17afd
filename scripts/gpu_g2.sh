# round-2 second-session round trip: GPU tests, the headline at K = 20, the other workloads, dataflow sweeps forced on for the
# single-scene workloads (RP_FLOW=2), hull build times. `gpurun --timeout 1500 -- 'bash scripts/gpu_g2.sh TAG'`
T=${1:-g2}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -12 > gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
summ() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', round(d['value']/1e6,2), round(d['ms_per_step'],3), round(d.get('e2e',{}).get('value',0)/1e6,2), {k:round(v['ms'],1) for k,v in d.get('kernels',{}).items()}, d.get('status_bits'), d.get('parity',{}).get('ok'))"; }
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err || tail -5 gpurun_out/${T}_bench20.err
summ gpurun_out/${T}_bench20.json
for WL in c2 c5 c3 pile; do
  timeout 400 python bench.py --workload $WL --warmup 3 --no-cpu > gpurun_out/${T}_$WL.json 2> gpurun_out/${T}_$WL.err || tail -5 gpurun_out/${T}_$WL.err
  summ gpurun_out/${T}_$WL.json
done
for WL in c3 pile; do
  RP_FLOW=2 timeout 400 python bench.py --workload $WL --warmup 3 --no-cpu > gpurun_out/${T}_${WL}_flow.json 2> gpurun_out/${T}_${WL}_flow.err || tail -5 gpurun_out/${T}_${WL}_flow.err
  summ gpurun_out/${T}_${WL}_flow.json
done
RP_FLOW=0 timeout 400 python bench.py --workload c5 --warmup 3 --no-cpu > gpurun_out/${T}_c5_noflow.json 2> gpurun_out/${T}_c5_noflow.err
summ gpurun_out/${T}_c5_noflow.json
timeout 120 python scripts/hull_build_time.py > gpurun_out/${T}_hulls.txt 2>&1; tail -25 gpurun_out/${T}_hulls.txt
