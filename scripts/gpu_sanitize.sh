# compute-sanitizer over a reduced run of every kernel family: `gpurun --timeout 2400 -- 'bash scripts/gpu_sanitize.sh TAG'`
T=${1:-san}
mkdir -p gpurun_out
for TOOL in memcheck racecheck synccheck; do
  ( time timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python scripts/sanitize_driver.py ) > gpurun_out/${T}_$TOOL.txt 2>&1
  echo "== $TOOL"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok |done|real" gpurun_out/${T}_$TOOL.txt | tail -16
done
