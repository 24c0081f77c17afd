mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r11_tests.txt
cat gpurun_out/r11_tests.txt
python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/r11_base.json 2> gpurun_out/r11_base.err
for v in pos3 man6 man3 epa6 epa12 gjk12 int6 int3; do
  RAWPHYS_B200_LIB=$PWD/raw-physics_b200/variants/lib_$v.so python bench.py --steps 60 --warmup 2 --no-cpu > gpurun_out/r11_$v.json 2>gpurun_out/r11_$v.err
done
for v in base pos3 man6 man3 epa6 epa12 gjk12 int6 int3; do
  python -c "
import json,sys; d=json.load(open('gpurun_out/r11_$v.json')); print('$v', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['status_bits'], {k:round(v['ms'],0) for k,v in d['kernels'].items()})"
done
for k in k_schedule k_broad_rows k_islands; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s 0 -c 1 -o gpurun_out/p11_$k -f python bench.py --ncu-frame 45 > gpurun_out/p11_$k.log 2>&1
done
