#!/bin/bash
# 1 GPU: compute-sanitizer over the kernels added in round 2 (tests/sanitizer_cases.py, SANITIZER_ROUND=2)
mkdir -p gpurun_out
rm -f gpurun_out/r02_sanitizer.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tests/sanitizer_cases.py  (SANITIZER_ROUND=2, 1x B200, round 2)" >> gpurun_out/r02_sanitizer.txt
  skip=""; if [ $tool = racecheck ]; then skip=1; fi
  SANITIZER_SKIP_WP=$skip SANITIZER_ROUND=2 timeout 900 compute-sanitizer --tool $tool python tests/sanitizer_cases.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  grep -E "^ok|SANITIZER_CASES_DONE|ERROR SUMMARY|RACECHECK SUMMARY|Traceback" gpurun_out/r02_sanitizer_$tool.log >> gpurun_out/r02_sanitizer.txt
  grep -E "Error|error|hazard" gpurun_out/r02_sanitizer_$tool.log | sort | uniq -c | sort -rn | head -8 >> gpurun_out/r02_sanitizer.txt
done
cat gpurun_out/r02_sanitizer.txt | cut -c1-300
grep -m1 -A12 "Barrier error" gpurun_out/r02_sanitizer_synccheck.log | cut -c1-300
