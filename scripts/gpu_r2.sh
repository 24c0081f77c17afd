# round-2 development round trip: `gpurun --timeout 900 -- 'bash scripts/gpu_r2.sh TAG [item ...]'`
# GPU tests (skipped when TAG starts with "nt"), then the bench at the driver's setting (K = 20: frames 40..59) and over the
# whole window (K = 60) for the product library, then K = 20 for every tuning item: VARIANT[@ENV=VAL[,ENV=VAL]] where VARIANT
# is a library under raw-physics_b200/variants/ (built by scripts/build_variants.py) or "-" for the product library
T=${1:-r2}; shift
mkdir -p gpurun_out
if [[ $T != nt* ]]; then
  ( time timeout 1200 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -12 > gpurun_out/${T}_tests.txt
  cat gpurun_out/${T}_tests.txt
fi
summ() { python -c "
import json,sys
d=json.load(open('$1')); print('$1', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d.get('e2e',{}).get('value',0)/1e6,1), {k:round(v['ms'],0) for k,v in d.get('kernels',{}).items()}, d.get('status_bits'), d.get('parity'))"; }
for K in 20 60; do
  timeout 300 python bench.py --steps $K --warmup 3 --no-cpu > gpurun_out/${T}_bench$K.json 2> gpurun_out/${T}_bench$K.err || tail -5 gpurun_out/${T}_bench$K.err
  summ gpurun_out/${T}_bench$K.json
done
for ITEM in "$@"; do
  V=${ITEM%%@*}; E=""; [[ $ITEM == *@* ]] && E=${ITEM#*@}
  N=$(echo $ITEM | tr '@=,' '___')
  LIB=""; [[ $V != "-" ]] && LIB="RAWPHYS_B200_LIB=$PWD/raw-physics_b200/variants/$V.so"
  env $LIB $(echo $E | tr ',' ' ') timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-parts 0 > gpurun_out/${T}_${N}.json 2> gpurun_out/${T}_${N}.err || tail -5 gpurun_out/${T}_${N}.err
  summ gpurun_out/${T}_${N}.json
done
