#!/bin/bash
# 1 GPU: four producer warps with deeper rounds and other register splits (variants 7, 8, 9) against variant 3
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ws_kernel.py -q -x -k "ws_4producers" > gpurun_out/r02_s23_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s23_pytest.log )
tail -n 4 gpurun_out/r02_s23_pytest.log
TOURNAMENT_VARIANTS=03789 timeout 300 build/ws_tournament -1 3 5 50 | tee gpurun_out/r02_s23_tournament.log
