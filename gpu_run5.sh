mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r5_tests.txt
cat gpurun_out/r5_tests.txt
python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/r5_base.json 2> gpurun_out/r5_base.err
tail -3 gpurun_out/r5_base.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r5_base.json')); print('base', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['status_bits'], {k:round(v['ms'],0) for k,v in d['kernels'].items()})"
date
for k in k_manifold k_pos_level k_integrate k_epa; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s 10 -c 1 -o gpurun_out/p5_$k -f python bench.py --ncu-frame 45 > gpurun_out/p5_$k.log 2>&1
  date
done
ls -la gpurun_out | grep p5
