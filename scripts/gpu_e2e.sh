# end-to-end leg with more sub-batches: `gpurun --timeout 900 -- 'bash scripts/gpu_e2e.sh TAG'`
T=${1:-e2e}
mkdir -p gpurun_out
for P in 4 8; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-parts $P > gpurun_out/${T}_p$P.json 2> gpurun_out/${T}_p$P.err || tail -5 gpurun_out/${T}_p$P.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_p$P.json')); print('P=$P', round(d['value']/1e6,1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step'],2), d['e2e']['api'][:60], 'other', round(d['e2e_other_form']['value']/1e6,1))"
done
