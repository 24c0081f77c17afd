# baseline of HEAD: GPU tests, bench line, ncu launch list, one ncu --set full pass over a whole substep of frame 40
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) 2>&1 | tail -15 > gpurun_out/r14_tests.txt
cat gpurun_out/r14_tests.txt
python bench.py --steps 60 --warmup 3 --hetero > gpurun_out/r14_bench.json 2> gpurun_out/r14_bench.err
tail -3 gpurun_out/r14_bench.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r14_bench.json')); print('base', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['status_bits'], {k:round(v['ms'],0) for k,v in d['kernels'].items()}); print(d.get('heterogeneous')); print(d.get('cpu_baseline'))"
python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/r14_ref.json 2> gpurun_out/r14_ref.err
cat gpurun_out/r14_ref.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r14_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r14_launches.log 2>&1
wc -l gpurun_out/r14_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_(integrate|cull|gjk|epa|manifold|pos_level|vel_level)' -c 27 -o gpurun_out/r14_substep -f python bench.py --ncu-frame 40 > gpurun_out/r14_ncu.log 2>&1
tail -3 gpurun_out/r14_ncu.log
ncu -i gpurun_out/r14_substep.ncu-rep --page raw --csv > gpurun_out/r14_substep.raw.csv 2>/dev/null
ls -la gpurun_out; du -sm gpurun_out
