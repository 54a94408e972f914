#!/bin/bash
# 1 GPU: whole GPU suite (incl. the new rhs / error tests), N=1 static vs dynamic item claiming
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_s7_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s7_pytest.log )
tail -n 30 gpurun_out/r02_s7_pytest.log
timeout 100 build/ws_tournament 3 3 5 50 | tee gpurun_out/r02_s7_static.log
EXADG_B200_DYNAMIC_ITEMS=1 timeout 100 build/ws_tournament 3 3 5 50 | tee gpurun_out/r02_s7_dynamic.log
