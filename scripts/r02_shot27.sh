#!/bin/bash
# 1 GPU: compute-sanitizer over the kernels added in round 2 (tests/sanitizer_cases.py, SANITIZER_ROUND=2)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tests/sanitizer_cases.py  (SANITIZER_ROUND=2, 1x B200, round 2)" >> gpurun_out/r02_sanitizer.txt
  SANITIZER_ROUND=2 timeout 900 compute-sanitizer --tool $tool python tests/sanitizer_cases.py 2>&1 | grep -E "^ok|SANITIZER_CASES_DONE|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|Traceback" | head -60 >> gpurun_out/r02_sanitizer.txt
done
tail -n 60 gpurun_out/r02_sanitizer.txt
