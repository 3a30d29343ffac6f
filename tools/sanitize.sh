#!/bin/bash
# compute-sanitizer over the small-shape driver, one tool at a time; logs -> gpurun_out/r2_sanitizer_<tool>.log
#   tools/sanitize.sh [tool ...]        (default: memcheck racecheck synccheck)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck}
# the same driver without a sanitizer: every kernel family at small shapes against the exact FFMA path
python tools/sanitize_driver.py > gpurun_out/r2_sanitize_driver_plain.log 2>&1; echo "plain driver exit code $?" | tee -a gpurun_out/r2_sanitize_driver_plain.log
tail -2 gpurun_out/r2_sanitize_driver_plain.log
for tool in $TOOLS; do
  log=gpurun_out/r2_sanitizer_${tool}.log
  echo "== compute-sanitizer --tool $tool" | tee "$log"
  timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 python tools/sanitize_driver.py >> "$log" 2>&1
  echo "exit code $?" | tee -a "$log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize_driver: done" "$log" | tail -8
done
