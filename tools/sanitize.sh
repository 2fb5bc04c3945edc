#!/bin/bash
# compute-sanitizer over every kernel instantiation on small batches (run under gpurun).
# Logs -> gpurun_out/sanitize_<tool>.log ; copy the summaries to profiles/.
mkdir -p gpurun_out
for tool in ${@:-racecheck synccheck memcheck}; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|WORST|hazard" gpurun_out/sanitize_$tool.log | head -8
  # the opt-in kernel that filters inside the synthesis kernel (spin-waits between its workers: racecheck only sees smem)
  AACFB_TNS_FUSED=1 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py tns_ar tns_ma > gpurun_out/sanitize_${tool}_fused.log 2>&1
  echo "== $tool (AACFB_TNS_FUSED=1) rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|WORST|hazard" gpurun_out/sanitize_${tool}_fused.log | head -8
done
