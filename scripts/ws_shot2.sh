#!/bin/bash
mkdir -p gpurun_out
timeout 35 build/ws_tournament -1 3 5 50 1 2 3 5 0 3 3 1 3 1 3 3 > gpurun_out/ws2_tournament.log 2>&1; echo "rc $?" >> gpurun_out/ws2_tournament.log
cat gpurun_out/ws2_tournament.log
