#!/bin/bash
# one GPU call: kernel tournament (parity + timing), ncu capture of the warp-specialised kernel, GPU test suite with that kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/ws_smi.txt 2>&1
timeout 60 build/ws_tournament 3 5 100 > gpurun_out/ws_tournament_96.log 2>&1; echo "rc96 $?" >> gpurun_out/ws_tournament_96.log
timeout 30 build/ws_tournament 3 4 50 > gpurun_out/ws_tournament_48.log 2>&1; echo "rc48 $?" >> gpurun_out/ws_tournament_48.log
timeout 70 env EXADG_B200_CART_KERNEL=ws python -m pytest tests -m gpu -x -q > gpurun_out/ws_pytest.log 2>&1; echo "rc $?" >> gpurun_out/ws_pytest.log
timeout 45 ncu --set full --import-source on --clock-control none -k regex:ws_kernel -c 1 -f -o gpurun_out/r01_cart_ws_k4_48 build/ws_tournament 3 4 3 1 > gpurun_out/ws_ncu.log 2>&1; echo "rc $?" >> gpurun_out/ws_ncu.log
tail -n 12 gpurun_out/ws_tournament_96.log gpurun_out/ws_tournament_48.log; tail -n 5 gpurun_out/ws_pytest.log; tail -n 3 gpurun_out/ws_ncu.log
