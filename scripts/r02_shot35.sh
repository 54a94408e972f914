#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:vmult_general_kernel -s 3 -c 1 -f -o gpurun_out/r02_general_k4_curved_64_v2 python bench.py --degree 4 --cells 64 --mesh curvilinear --steps 2 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain > gpurun_out/r02_s35_ncu.log 2>&1
tail -2 gpurun_out/r02_s35_ncu.log
