#!/bin/bash
# warp-private k=4 kernel: first GPU run.  parity tests of all variants, tournament at 48^3 and 96^3, ncu of variant 4
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_ws_kernel.py -x -q > gpurun_out/r02_s2_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s2_pytest.log )
tail -n 6 gpurun_out/r02_s2_pytest.log
TOURNAMENT_SKIP2=1 timeout 120 build/ws_tournament -1 3 4 50 3 5 50 > gpurun_out/r02_s2_tournament.log 2>&1; echo "rc $?" >> gpurun_out/r02_s2_tournament.log
cat gpurun_out/r02_s2_tournament.log
timeout 200 ncu --set full --import-source on --clock-control none -k regex:wp_kernel -s 3 -c 1 -f -o gpurun_out/r02_cart_wp_k4_96 build/ws_tournament 4 3 5 3 > gpurun_out/r02_ncu_wp.log 2>&1
tail -n 3 gpurun_out/r02_ncu_wp.log
