#!/bin/bash
# ncu capture of the k=4 kernels on the bench workload (96^3 cells): launch durations and one full capture of the default kernel.
#   gpurun --timeout 300 -- 'bash scripts/ncu_ws.sh'      (results in gpurun_out/, summaries go to profiles/)
mkdir -p gpurun_out
timeout 120 ncu --set full --import-source on --clock-control none -k regex:ws_kernel -c 1 -f -o gpurun_out/cart_ws_k4_96 build/ws_tournament 1 3 5 3 > gpurun_out/ncu_ws.log 2>&1
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_ws.csv build/ws_tournament -1 3 5 3 > gpurun_out/ncu_ws_launches.log 2>&1
tail -n 3 gpurun_out/ncu_ws.log
