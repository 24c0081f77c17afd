# single-precision fast mode: GPU tests, then bench lines of the f32 build (and occupancy variants of it) on the headline and the
# single-scene workloads. `gpurun --timeout 1500 -- 'bash scripts/gpu_f32.sh TAG [variant ...]'`
T=${1:-f32}; shift
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -12 > gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
summ() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', d.get('dtype'), round(d['value']/1e6,2), round(d['ms_per_step'],3), 'window', round(d['window']['value']/1e6,1), 'e2e', round(d.get('e2e',{}).get('value',0)/1e6,2), {k:round(v['ms'],1) for k,v in d.get('kernels',{}).items()}, d.get('status_bits'), d.get('parity'))"; }
timeout 300 python bench.py --precision f32 --steps 20 --warmup 3 --no-cpu > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err || tail -5 gpurun_out/${T}_bench20.err
summ gpurun_out/${T}_bench20.json
for V in "$@"; do
  RP_F32_LIB=$PWD/raw-physics_b200/variants/$V.so timeout 300 python bench.py --precision f32 --steps 20 --warmup 3 --no-cpu --e2e-parts 0 > gpurun_out/${T}_$V.json 2> gpurun_out/${T}_$V.err || tail -5 gpurun_out/${T}_$V.err
  summ gpurun_out/${T}_$V.json
done
for WL in c2 c3 pile c5; do
  timeout 400 python bench.py --precision f32 --workload $WL --warmup 3 --no-cpu > gpurun_out/${T}_$WL.json 2> gpurun_out/${T}_$WL.err || tail -5 gpurun_out/${T}_$WL.err
  summ gpurun_out/${T}_$WL.json
done
