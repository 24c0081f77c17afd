mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) 2>&1 | tail -15 > gpurun_out/r18_tests.txt
cat gpurun_out/r18_tests.txt
python bench.py --steps 60 --warmup 3 --no-cpu --hetero > gpurun_out/r18_bench.json 2> gpurun_out/r18_bench.err
tail -3 gpurun_out/r18_bench.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r18_bench.json')); print('base', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['status_bits'], {k:round(v['ms'],0) for k,v in d['kernels'].items()}); print(d.get('heterogeneous'))"
for v in epa8 epa9 man5 man6 int8 gjk10 gjk12; do
  RAWPHYS_B200_LIB=$PWD/raw-physics_b200/variants/lib_$v.so python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/var_$v.json 2>gpurun_out/var_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/var_$v.json')); print('$v', round(d['value']/1e6,1), round(d['ms_per_step'],2), {k:round(v['ms'],0) for k,v in d['kernels'].items() if v['ms']>60})"
done
