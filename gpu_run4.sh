mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r4_tests.txt
cat gpurun_out/r4_tests.txt
python bench.py --steps 60 --warmup 3 --no-cpu > gpurun_out/r4_base.json 2> gpurun_out/r4_base.err
tail -3 gpurun_out/r4_base.err
for v in base; do
  python -c "
import json,sys; d=json.load(open('gpurun_out/r4_$v.json')); print('$v', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), d['status_bits'], {k:round(v['ms'],0) for k,v in d['kernels'].items()})"
done
