mkdir -p gpurun_out
python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/r20_bench.json 2> gpurun_out/r20_bench.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r20_bench.json')); print('base', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['status_bits'], {k:round(v['ms'],0) for k,v in d['kernels'].items()})"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_(gjk|epa|manifold)' -c 3 -o gpurun_out/r20_hetero -f python bench.py --ncu-frame 40 --hetero > gpurun_out/r20_ncu.log 2>&1
tail -3 gpurun_out/r20_ncu.log
ncu -i gpurun_out/r20_hetero.ncu-rep --page raw --csv > gpurun_out/r20_hetero.raw.csv 2>/dev/null
ncu -i gpurun_out/r20_hetero.ncu-rep --page source --csv -k regex:k_epa > gpurun_out/r20_hetero_epa.source.csv 2>/dev/null
ls -la gpurun_out | tail -5
