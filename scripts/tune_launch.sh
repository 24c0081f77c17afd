run() { L=$1; shift
  timeout 300 python bench.py "$@" --no-cpu --e2e-parts 0 --no-extras > gpurun_out/g3_$L.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/g3_$L.json')); print('$L', round(d['value']/1e6,1), round(d['ms_per_step'],3), d['parity_checked'])"; }
run c5 --workload c5
run c2 --workload c2
run c3 --workload c3
run w20 --steps 20
run w60 --steps 60
run pile --workload pile
