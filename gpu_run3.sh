mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r3_tests.txt
cat gpurun_out/r3_tests.txt
python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/r3_base.json 2> gpurun_out/r3_base.err
for v in sorted pos3 gjk6 epa6 int5; do
  RAWPHYS_B200_LIB=$PWD/raw-physics_b200/variants/lib_$v.so python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/r3_$v.json 2>gpurun_out/r3_$v.err
done
for v in base sorted pos3 gjk6 epa6 int5; do
  python -c "
import json,sys; d=json.load(open('gpurun_out/r3_$v.json')); print('$v', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['status_bits'], {k:round(v['ms'],0) for k,v in d['kernels'].items()})"
done
