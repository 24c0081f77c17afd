mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) 2>&1 | tail -8 > gpurun_out/r33_tests.txt
cat gpurun_out/r33_tests.txt
python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/r33_bench.json 2> gpurun_out/r33_bench.err
tail -3 gpurun_out/r33_bench.err
python -c "
import json,sys
d=json.load(open('gpurun_out/r33_bench.json')); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],0) for k,v in d['kernels'].items()})"
for v in vel4 vel2; do
  RAWPHYS_B200_LIB=$PWD/raw-physics_b200/variants/lib_$v.so python bench.py --steps 60 --warmup 3 --no-cpu --e2e-parts 1 > gpurun_out/var_$v.json 2>gpurun_out/var_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/var_$v.json')); print('$v', round(d['value']/1e6,1), round(d['ms_per_step'],2), {k:round(v['ms'],0) for k,v in d['kernels'].items() if v['ms']>30})"
done
