mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_coloured.py -m gpu -q -s ) > gpurun_out/r24_tests_col.txt 2>&1
grep -v "^$" gpurun_out/r24_tests_col.txt | tail -30
