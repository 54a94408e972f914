#!/bin/bash
mkdir -p gpurun_out
for v in 3 6 4; do
  echo "== synccheck variant $v"
  timeout 300 compute-sanitizer --tool synccheck --print-limit 2 python scripts/sync_case.py $v 2>&1 | grep -E "^ok|ERROR SUMMARY|Barrier error|    at |Device Frame" | head -8
done
TOURNAMENT_VARIANTS=03 timeout 200 build/ws_tournament -1 3 5 50
