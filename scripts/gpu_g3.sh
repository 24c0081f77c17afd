# GPU tests + the short sanitizer passes (memcheck, synccheck) over the reduced driver: `gpurun --timeout 1500 -- 'bash scripts/gpu_g3.sh TAG'`
T=${1:-g3}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -12 > gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
for TOOL in memcheck synccheck; do
  ( time timeout 600 compute-sanitizer --tool $TOOL --print-limit 20 python scripts/sanitize_driver.py ) > gpurun_out/${T}_$TOOL.txt 2>&1
  echo "== $TOOL"; grep -E "ERROR SUMMARY|ok |done|real|Error|error" gpurun_out/${T}_$TOOL.txt | tail -24
done
