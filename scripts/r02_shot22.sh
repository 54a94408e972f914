#!/bin/bash
# 1 GPU: staged variant (6) of the warp-specialised kernel: parity tests, then the tournament 0 / 3 / 6 at 96^3
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ws_kernel.py -q -x -k "ws_staged or ws_4producers" > gpurun_out/r02_s22_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s22_pytest.log )
tail -n 6 gpurun_out/r02_s22_pytest.log
TOURNAMENT_VARIANTS=036 timeout 200 build/ws_tournament -1 3 5 50 | tee gpurun_out/r02_s22_tournament.log
