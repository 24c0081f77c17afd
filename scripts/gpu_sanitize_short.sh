# compute-sanitizer memcheck + synccheck over the reduced driver (racecheck takes 19 minutes: scripts/gpu_sanitize.sh)
T=${1:-sanshort}
mkdir -p gpurun_out
for TOOL in memcheck synccheck; do
  ( time timeout 200 compute-sanitizer --tool $TOOL --print-limit 20 python scripts/sanitize_driver.py ) > gpurun_out/${T}_$TOOL.txt 2>&1
  echo "== $TOOL"; grep -E "ERROR SUMMARY|done|real" gpurun_out/${T}_$TOOL.txt | tail -4
done
