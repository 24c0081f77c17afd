mkdir -p gpurun_out
for v in c8 c32 c64; do
  RAWPHYS_B200_LIB=$PWD/raw-physics_b200/variants/lib_$v.so python bench.py --steps 60 --warmup 3 --no-cpu --e2e-parts 1 > gpurun_out/var_$v.json 2>gpurun_out/var_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/var_$v.json')); print('$v', round(d['value']/1e6,1), round(d['ms_per_step'],2), {k:round(v['ms'],0) for k,v in d['kernels'].items() if v['ms']>30})"
done
