run() { L=$1; V=$2; shift; shift
  LIB=""; [[ $V != "-" ]] && LIB="RAWPHYS_B200_LIB=$PWD/raw-physics_b200/variants/$V.so"
  env $LIB timeout 300 python bench.py "$@" --no-cpu --e2e-parts 0 --no-extras > gpurun_out/g5_$L.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/g5_$L.json')); print('$L', round(d['value']/1e6,1), round(d['ms_per_step'],3), d['parity_checked'])"; }
for V in - big1 big2; do
  run w20_$V $V --steps 20
  run c3_$V $V --workload c3
  run c5_$V $V --workload c5
  run c2_$V $V --workload c2
done
