#!/bin/bash
# compute-sanitizer over every kernel instantiation on small batches (run under gpurun).
# Logs -> gpurun_out/sanitize_<tool>.log ; copy the summaries to profiles/.
mkdir -p gpurun_out
for tool in ${@:-racecheck synccheck memcheck}; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|WORST|hazard" gpurun_out/sanitize_$tool.log | head -8
done
